python tools/e2e_pipelined.py 256 20
NCTX=3 python tools/e2e_pipelined.py 256 20
python tools/e2e_pipelined.py 128 40
python tools/e2e_pipelined.py 256 10 pageable
for T in 4 12 16; do ORBX_STAGE_THREADS=$T NCTX=1 python tools/e2e_pipelined.py 256 10 pageable; done
