TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29561 tools/pcie_bw.py 2>/dev/null | tail -1 > gpurun_out/m2_pcie_bw_2gpu.json; cat gpurun_out/m2_pcie_bw_2gpu.json
timeout 300 $TR --master-port 29562 bench.py --gpus 2 --config c4 --warmup 5 > gpurun_out/m2_bench_c4_2gpu.json 2> gpurun_out/m2_c4.err; cut -c1-300 gpurun_out/m2_bench_c4_2gpu.json
timeout 300 $TR --master-port 29563 bench.py --gpus 2 --config c2 --steps 200 --warmup 5 > gpurun_out/m2_bench_c2_2gpu.json 2> gpurun_out/m2_c2.err; cut -c1-300 gpurun_out/m2_bench_c2_2gpu.json
