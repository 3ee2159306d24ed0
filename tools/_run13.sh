TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29571 tools/pcie_bw.py 2>/dev/null | tail -1 > gpurun_out/m8_pcie_bw_8gpu.json; cat gpurun_out/m8_pcie_bw_8gpu.json
timeout 400 $TR --master-port 29572 bench.py --gpus 8 --config c4 --warmup 5 > gpurun_out/m8_bench_c4_8gpu.json 2> gpurun_out/m8_c4.err; cut -c1-300 gpurun_out/m8_bench_c4_8gpu.json
timeout 500 $TR --master-port 29573 bench.py --gpus 8 --config c5 --warmup 3 > gpurun_out/m8_bench_c5_8gpu.json 2> gpurun_out/m8_c5.err; cut -c1-300 gpurun_out/m8_bench_c5_8gpu.json
timeout 300 $TR --master-port 29574 bench.py --gpus 8 --config c2 --steps 200 --warmup 5 > gpurun_out/m8_bench_c2_8gpu.json 2> gpurun_out/m8_c2.err; cut -c1-300 gpurun_out/m8_bench_c2_8gpu.json
ORBX_HOST_LANES=4 timeout 300 $TR --master-port 29575 bench.py --gpus 8 --config c2 --steps 50 --warmup 5 > gpurun_out/m8_bench_c2_8gpu_lanes4.json 2> gpurun_out/m8_c2b.err
ORBX_HOST_LANES=12 timeout 300 $TR --master-port 29576 bench.py --gpus 8 --config c2 --steps 50 --warmup 5 > gpurun_out/m8_bench_c2_8gpu_lanes12.json 2> gpurun_out/m8_c2c.err
python - <<'P'
import json
for f in ['m8_bench_c2_8gpu','m8_bench_c2_8gpu_lanes4','m8_bench_c2_8gpu_lanes12','m8_bench_c4_8gpu','m8_bench_c5_8gpu']:
    try:
        d=json.load(open(f'gpurun_out/{f}.json')); print(f, round(d['value']), round(d['e2e']['value']), d['parity'].get('keypoint_record_mismatches'), d['parity'].get('match_record_mismatches'))
    except Exception as e: print(f, 'failed', e)
P
