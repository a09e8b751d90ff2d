# weak + strong scaling bench at G GPUs (+ the multi-GPU parity worker), each command under its own timeout
G=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29671 tests/multi_gpu_worker.py > gpurun_out/r2_multi_worker_n$G.log 2>&1; grep -E "PARITY" gpurun_out/r2_multi_worker_n$G.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29672 bench.py --gpus $G --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2_scale_n${G}_weak.json 2> gpurun_out/r2_scale_n${G}_weak.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29673 bench.py --gpus $G --steps 30 --warmup 3 --no-cpu-baseline --scaling strong > gpurun_out/r2_scale_n${G}_strong.json 2> gpurun_out/r2_scale_n${G}_strong.err
python tools/show_bench.py gpurun_out/r2_scale_n${G}_weak.json gpurun_out/r2_scale_n${G}_strong.json
