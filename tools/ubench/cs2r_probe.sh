#!/usr/bin/env bash
# Builds tools/ubench/cs2r_probe.cu in its variants and prints, per variant, the closest CS2R -> reader distance found by
# tools/sass_hazard_scan.py; run the binaries on a B200 (`gpurun -- bash tools/ubench/cs2r_probe.sh run`).
cd "$(dirname "$0")"
for v in "O3:" "O1:-Xptxas -O1" "uncond:-DPROBE_UNCONDITIONAL" "uncondO1:-DPROBE_UNCONDITIONAL -Xptxas -O1"; do
    name=${v%%:*}; flags=${v#*:}
    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a $flags -DPROBE_NAME="\"$name\"" cs2r_probe.cu -o cs2r_probe_$name || exit 1
    echo "== $name: $(python ../sass_hazard_scan.py cs2r_probe_$name --min-cycles 7 2>&1 | tail -1)"
    if [ "$1" = run ]; then ./cs2r_probe_$name | tail -12; fi
done
