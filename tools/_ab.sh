cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python tools/normals_bench.py 10000000 5 2>&1 | tail -1
python tools/ndt_bench.py 2>&1 | tail -3 | cut -c1-400
