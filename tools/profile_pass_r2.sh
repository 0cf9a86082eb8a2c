#!/bin/bash
# Round-2 profile pass (one GPU call): plain runs first (each must exit 0), then ncu launch lists and full captures.
# The captures are summarised ON the box (tools/ncu_summary.py) and only reports below 8 MB travel back: gpurun_out/ is
# capped at 64 MiB per call.
# usage (under gpurun): bash tools/profile_pass_r2.sh <tag> [sections...]   sections: c1 c1b c2 c3 c4 c5 vx (default: all)
set -u
TAG=${1:-r2}; shift
SECT=${*:-c1 c1b c2 c3 c4 c5 vx}
OUT=gpurun_out
TMP=/tmp/b2prof; mkdir -p $TMP
has() { [[ " $SECT " == *" $1 "* ]]; }
# cap <name> <kernel regex> <skip> <count> <command...>
cap() {
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  "$@" > /dev/null 2>&1 || { echo "plain run failed: $*"; return; }
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$rx" -s $skip -c $cnt -f -o $TMP/${TAG}_prof_$name "$@" > $OUT/${TAG}_ncu_$name.log 2>&1
  python tools/ncu_summary.py $TMP/${TAG}_prof_$name.ncu-rep > $OUT/${TAG}_ncu_$name.txt 2>/dev/null
  local sz=$(stat -c %s $TMP/${TAG}_prof_$name.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 0 ] && [ "$sz" -lt 8000000 ]; then cp $TMP/${TAG}_prof_$name.ncu-rep $OUT/; fi
  tail -c 300 $OUT/${TAG}_ncu_$name.log > $OUT/${TAG}_ncu_$name.log.tail; mv $OUT/${TAG}_ncu_$name.log.tail $OUT/${TAG}_ncu_$name.log
}
# list <name> <count> <command...>
list() {
  local name=$1 cnt=$2; shift 2
  "$@" > $OUT/${TAG}_plain_$name.log 2>&1 || { echo "plain run failed: $*"; return; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c $cnt --csv --log-file $OUT/${TAG}_launches_$name.csv "$@" > /dev/null 2>&1
}
C1="python bench.py --steps 2 --warmup 3 --batch 0 --skip-registration"
C1B="python tools/prof_s2m_batch.py"
C2="python tools/frontend_bench.py"
C3="python tools/ndt_bench.py"
C4="python tools/gicp_bench.py --skip-c5"
C5="python tools/gicp_bench.py --skip-c4 --c5-points ${C5_POINTS:-50000000}"
VX="python tools/voxel_bench.py"
if has c1; then
  list c1 300 $C1
  cap s2m "k_s2m_iteration<\(int\)16" 6 2 $C1
  cap gridbuild "k_grid_build_dev" 6 2 $C1
fi
if has c1b; then cap s2m_batched "k_s2m_iteration<\(int\)1, " 4 4 $C1B; fi
if has c2; then
  list c2 200 $C2
  cap scan "k_scan_" 100 10 $C2
fi
if has c3; then
  list c3 300 $C3
  cap ndt "k_ndt_derivatives|k_ndt_fitness" 14 3 $C3
fi
if has c4; then
  list c4 400 $C4
  cap c4 "k_gicp_linearize|k_normals" 8 4 $C4
fi
if has vx; then
  list vx 200 $VX
  cap vx "k_vx|k_rs_|k_scan_apply|k_scan_tile" 30 30 $VX
fi
if has c5; then
  list c5 200 $C5
  cap gicp "k_gicp_linearize" 0 6 $C5
fi
du -sh $OUT; ls -la $OUT | grep ${TAG}_ | awk '{print $5, $9}'
