#!/bin/bash
# usage (under gpurun): tools/profile_one.sh <tag> <kernel-regex> [workload] [variant]   -> gpurun_out/<tag>.ncu-rep
tag=$1; pat=$2; wl=${3:-c3}; lib=${4:-default}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$pat -s 3 -c 1 -f -o gpurun_out/$tag python tools/ab.py --workload $wl --steps 2 --warmup 2 --no-parity $lib > gpurun_out/$tag.log 2>&1
ls -la gpurun_out/$tag.ncu-rep
