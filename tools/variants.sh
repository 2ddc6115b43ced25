#!/bin/bash
# developer helper: per-kernel times of the 1024-primitive scene for several builds of the library
for v in "$@"; do
    if [ "$v" = main ]; then unset SDM_LIB; else export SDM_LIB=$PWD/build/variants/$v.so; fi
    echo "#### variant $v"
    python tools/perf_probe.py many1024_1024 2>&1 | grep -v "k_bitscan\|k_weld\|k_emit\|k_clear\|clears\|k_init\|start"
done
