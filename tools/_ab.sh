cd /root/repo
timeout 900 python -m pytest tests/test_gpu_ndt.py tests/test_real_data.py -x -q -m gpu 2>&1 | tail -2
python tools/ndt_bench.py 2>&1 | tail -3 | cut -c1-500
