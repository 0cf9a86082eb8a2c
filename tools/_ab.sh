cd /root/repo
timeout 900 python -m pytest tests/test_gpu_gicp.py tests/test_real_data.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
python tools/gicp_bench.py 2>&1 | grep -o '"evaluation_ms": [^]]*]\|"gpu_ms": [0-9.]*\|"mpts_per_s": [0-9.]*\|"t_err": [0-9.e-]*\|"index": [0-9.]*'
