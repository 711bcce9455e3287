#!/usr/bin/env bash
# guarded.sh SECONDS command...  -- run one risky GPU step under a hard timeout and STOP the whole gpurun script (exit 99)
# if it timed out or if the device does not answer afterwards, instead of letting every later step run into its own
# timeout (round 1 lost 6 GPU-minutes that way: tools/experiments/README.md).
#   usage inside a gpurun command:   tools/experiments/guarded.sh 40 python bench.py --steps 10 --warmup 3 || exit 1
limit=$1; shift
timeout -k 5 "$limit" "$@"
rc=$?
if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then
    echo "guarded: '$*' exceeded ${limit}s -- stopping" >&2
    exit 99
fi
if ! timeout 20 nvidia-smi --query-gpu=name --format=csv,noheader > /dev/null 2>&1; then
    echo "guarded: device does not answer after '$*' -- stopping" >&2
    exit 99
fi
exit $rc
