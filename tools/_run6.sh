python tools/step_time.py 256 50
ORBX_OVERLAP=0 python tools/step_time.py 256 50
ORBX_OVERLAP=0 ORBX_TAIL_PX=0 python tools/step_time.py 256 50
ORBX_TAIL_PX=0 python tools/step_time.py 256 50
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
