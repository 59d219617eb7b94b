#!/bin/bash
# tools/ncu_capture.sh TAG KERNEL_REGEX [CELLS]   -- one `ncu --set full` capture of the first matching launch (after the
# program has run once without ncu), report -> gpurun_out/prof_TAG.ncu-rep.  Run under gpurun (1 GPU).
TAG=$1; KRE=$2; CELLS=${3:-18944}
python bench.py --no-e2e --no-cpu --steps 1 --warmup 1 --cells $CELLS > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:"$KRE" -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --no-e2e --no-cpu --steps 1 --warmup 1 --cells $CELLS > gpurun_out/ncu_$TAG.log 2>&1
