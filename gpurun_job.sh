cd en234_fea_b200/csrc
for v in c9o2 c3o4 c9o3; do cp liben234gpu_$v.so liben234gpu.so; echo $v; (cd ../..; python tools/mf_times.py 128 0.0; python tools/mf_times.py 64 0.0); done
