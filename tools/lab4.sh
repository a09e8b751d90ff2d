python -m pytest tests -m gpu -q -k "bfs or graph or sampled or config1 or rank" 2>&1 | tail -4
python tools/bfs_lab.py 2>&1 | tail -2
GM_BFS_V1=1 python tools/bfs_lab.py 2>&1 | tail -1
L=$PWD/matrix-manifolds_b200/lib
for v in 1 2 3; do GM_B200_LIB=$L/libgm_b200_hop$v.so python tools/sampled_lab.py 2>&1 | tail -1; done
GM_B200_LIB=$L/libgm_b200_hop3.so ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:spd_pair_stream -s 15 -c 2 python tools/sampled_lab.py 2>&1 | grep -E "dram__bytes_read|gpu__time" | tail -4
