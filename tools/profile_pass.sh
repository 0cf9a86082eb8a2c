#!/bin/bash
# One GPU call: plain runs first (each must exit 0), then the ncu launch lists and full captures quoted in profiles/README.md.
# usage (under gpurun): bash tools/profile_pass.sh <tag>
set -u
TAG=${1:-r1}
OUT=gpurun_out
C1="python bench.py --steps 2 --warmup 3 --batch 16 --skip-registration"
C5="python tools/gicp_bench.py --skip-c4 --c5-points ${C5_POINTS:-50000000}"
C4="python tools/gicp_bench.py --skip-c5"
C3="python tools/ndt_bench.py"
$C1 > $OUT/${TAG}_plain_c1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_c1.csv $C1 > $OUT/${TAG}_ncu_c1_list.log 2>&1
$C1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_s2m_iteration -s 8 -c 3 -o $OUT/${TAG}_prof_s2m $C1 > $OUT/${TAG}_ncu_s2m.log 2>&1
$C3 > $OUT/${TAG}_plain_c3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_c3.csv $C3 > $OUT/${TAG}_ncu_c3_list.log 2>&1
$C3 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_ndt_derivatives -s 14 -c 2 -o $OUT/${TAG}_prof_ndt $C3 > $OUT/${TAG}_ncu_ndt.log 2>&1
$C4 > $OUT/${TAG}_plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_normals -s 4 -c 1 -o $OUT/${TAG}_prof_normals $C4 > $OUT/${TAG}_ncu_normals.log 2>&1
$C5 > $OUT/${TAG}_plain_c5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches_c5.csv $C5 > $OUT/${TAG}_ncu_c5_list.log 2>&1
$C5 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gicp_linearize -c 6 -o $OUT/${TAG}_prof_gicp $C5 > $OUT/${TAG}_ncu_gicp.log 2>&1
ls -la $OUT | grep ${TAG}_
