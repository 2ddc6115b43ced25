#!/bin/bash
# one GPU: the tail kernel on the slow frames of configs[4]; tail parity tests; with "frame": that frame against the reference kernels at full size (6 min)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/c5_frame_probe.py 254 329 100 > gpurun_out/y_probe.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/y_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/y_pytest.log
if [ "$1" = "frame" ]; then
SDM_SLOW_TESTS=1 timeout 900 python -m pytest tests/test_gpu_ref_oracle.py -m gpu -q -x -k "fullsize and 4.2" > gpurun_out/y_pytest_frame.log 2>&1; echo "pytest exit $?" >> gpurun_out/y_pytest_frame.log
fi
tail -7 gpurun_out/y_probe.log; tail -3 gpurun_out/y_pytest.log; [ "$1" = "frame" ] && tail -5 gpurun_out/y_pytest_frame.log
exit 0
