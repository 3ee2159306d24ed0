TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29551 tools/pcie_bw.py 2>/dev/null | tail -1 > gpurun_out/m4_pcie_bw_4gpu.json; cat gpurun_out/m4_pcie_bw_4gpu.json
timeout 300 $TR --master-port 29552 bench.py --gpus 4 --config c4 --warmup 5 > gpurun_out/m4_bench_c4_4gpu.json 2> gpurun_out/m4_c4.err; cut -c1-300 gpurun_out/m4_bench_c4_4gpu.json
timeout 300 $TR --master-port 29553 bench.py --gpus 4 --config c2 --steps 200 --warmup 5 > gpurun_out/m4_bench_c2_4gpu.json 2> gpurun_out/m4_c2.err; cut -c1-300 gpurun_out/m4_bench_c2_4gpu.json
