#!/bin/bash
# Round-2 profile pass (one GPU call): plain runs first (each must exit 0), then ncu launch lists and full captures.
# usage (under gpurun): bash tools/profile_pass_r2.sh <tag> [sections...]   sections: c1 c1b c2 c3 c4 c5 vx (default: all)
set -u
TAG=${1:-r2}; shift
SECT=${*:-c1 c1b c2 c3 c4 c5 vx}
OUT=gpurun_out
has() { [[ " $SECT " == *" $1 "* ]]; }
C1="python bench.py --steps 2 --warmup 3 --batch 0 --skip-registration"
C1B="python tools/prof_s2m_batch.py"
C2="python tools/frontend_bench.py"
C3="python tools/ndt_bench.py"
C4="python tools/gicp_bench.py --skip-c5"
C5="python tools/gicp_bench.py --skip-c4 --c5-points ${C5_POINTS:-50000000}"
VX="python tools/voxel_bench.py"
if has c1; then
$C1 > $OUT/${TAG}_plain_c1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_c1.csv $C1 > $OUT/${TAG}_ncu_c1_list.log 2>&1
$C1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_s2m_iteration<\(int\)16" -s 6 -c 2 -o $OUT/${TAG}_prof_s2m $C1 > $OUT/${TAG}_ncu_s2m.log 2>&1
$C1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_grid_build_dev -s 6 -c 2 -o $OUT/${TAG}_prof_gridbuild $C1 > $OUT/${TAG}_ncu_gridbuild.log 2>&1
fi
if has c1b; then
$C1B > $OUT/${TAG}_plain_c1b.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_s2m_iteration<\(int\)1, " -s 4 -c 4 -o $OUT/${TAG}_prof_s2m_batched $C1B > $OUT/${TAG}_ncu_s2mb.log 2>&1
fi
if has c2; then
$C2 > $OUT/${TAG}_plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches_c2.csv $C2 > $OUT/${TAG}_ncu_c2_list.log 2>&1
$C2 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:k_scan_|k_vx_|k_rs_" -s 60 -c 30 -o $OUT/${TAG}_prof_scan $C2 > $OUT/${TAG}_ncu_scan.log 2>&1
fi
if has c3; then
$C3 > $OUT/${TAG}_plain_c3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_c3.csv $C3 > $OUT/${TAG}_ncu_c3_list.log 2>&1
$C3 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:k_ndt_derivatives|k_ndt_fitness" -s 14 -c 3 -o $OUT/${TAG}_prof_ndt $C3 > $OUT/${TAG}_ncu_ndt.log 2>&1
fi
if has c4; then
$C4 > $OUT/${TAG}_plain_c4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_c4.csv $C4 > $OUT/${TAG}_ncu_c4_list.log 2>&1
$C4 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:k_gicp_linearize|k_normals" -s 8 -c 6 -o $OUT/${TAG}_prof_c4 $C4 > $OUT/${TAG}_ncu_c4.log 2>&1
fi
if has vx; then
$VX > $OUT/${TAG}_plain_vx.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches_vx.csv $VX > $OUT/${TAG}_ncu_vx_list.log 2>&1
$VX > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:k_vx_|k_rs_|k_scan_apply|k_scan_tile" -s 40 -c 40 -o $OUT/${TAG}_prof_vx $VX > $OUT/${TAG}_ncu_vx.log 2>&1
fi
if has c5; then
$C5 > $OUT/${TAG}_plain_c5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches_c5.csv $C5 > $OUT/${TAG}_ncu_c5_list.log 2>&1
$C5 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gicp_linearize -c 6 -o $OUT/${TAG}_prof_gicp $C5 > $OUT/${TAG}_ncu_gicp.log 2>&1
fi
ls -la $OUT | grep ${TAG}_ | awk '{print $5, $9}'
