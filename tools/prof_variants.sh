#!/bin/bash
# developer tool (GPU box): C1 iteration breakdown + e2e step for the default library and every variants/*.so, A/B on one box
cd "$(dirname "$0")/.."
run() {
  echo "=== $1"
  python tools/prof_iter.py 2>&1 | tail -5
  python tools/prof_c1_e2e.py 2>&1 | tail -1
}
run default
for so in multi_sensor_slam_tookit_b200/variants/libb2reg_*.so; do
  [ -e "$so" ] || continue
  B2_LIB=$PWD/$so run "$so"
done
echo "=== default again"; python tools/prof_c1_e2e.py 2>&1 | tail -1
echo "=== timeline (default)"
B2_S2M_TIMELINE=1 python tools/prof_c1_e2e.py 2>&1 | tail -3
