"""Developer tool: build kernel-experiment variants of libb2reg.so (variants/libb2reg_<name>.so, selected with B2_LIB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_sensor_slam_tookit_b200 import build
V = {}
for arg in sys.argv[1:]:
    name, _, flags = arg.partition("=")
    V[name] = flags.split(",") if flags else []
for name, flags in V.items():
    print(name, flags, build.build_lib(variant=name, variant_flags=flags))
