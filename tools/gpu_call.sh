python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02w_bench_2gpu.json 2> gpurun_out/r02w_bench_2gpu.err
echo "rc=$?"; tail -2 gpurun_out/r02w_bench_2gpu.err; python tools/show_bench.py gpurun_out/r02w_bench_2gpu.json 2>/dev/null | head -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02w_ref_2gpu.json 2> gpurun_out/r02w_ref_2gpu.err
echo "rc=$?"; tail -2 gpurun_out/r02w_ref_2gpu.err; cat gpurun_out/r02w_ref_2gpu.json | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 --workload cfg3 > gpurun_out/r02w_bench_cfg3_2gpu.json 2> gpurun_out/r02w_bench_cfg3_2gpu.err
echo "rc=$?"; python tools/show_bench.py gpurun_out/r02w_bench_cfg3_2gpu.json 2>/dev/null | head -2
