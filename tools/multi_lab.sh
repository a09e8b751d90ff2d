# exchange A/B at G GPUs: pipelined with K chunks vs the single fused kernel (device-timed leg only)
G=${1:-4}
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $G --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_x_n${G}_$tag.json 2> gpurun_out/r2_x_n${G}_$tag.err; python tools/show_bench.py gpurun_out/r2_x_n${G}_$tag.json | cut -c1-330; }
run fused GM_PEER_PIPELINE=0
run k2 GM_PEER_CHUNKS=2
run k4 GM_PEER_CHUNKS=4
run k8 GM_PEER_CHUNKS=8
