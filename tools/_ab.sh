cd /root/repo
timeout 900 python -m pytest tests/test_gpu_gicp.py tests/test_real_data.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
for v in default nogc; do
  echo "=== $v"
  if [ $v = default ]; then unset B2_LIB; else export B2_LIB=$PWD/multi_sensor_slam_tookit_b200/variants/libb2reg_$v.so; fi
  python tools/gicp_bench.py 2>&1 | grep -o '"normals_ms": [0-9.]*\|"gpu_ms": [0-9.]*\|"mpts_per_s": [0-9.]*\|"e2e_s": [0-9.]*\|"ms_per_pair[a-z_]*": [0-9.]*'
done
