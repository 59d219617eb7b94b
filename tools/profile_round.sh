#!/bin/bash
# tools/profile_round.sh TAG   -- run under gpurun on ONE GPU.  Leaves in gpurun_out/:
#   plain_TAG.json          the bench line of the profiled command, run WITHOUT ncu first (must exit 0)
#   launches_TAG.csv        ncu launch list (gpu__time_duration.sum per launch, --clock-control none) of the same command,
#                           restricted to this library's kernels (the synthetic-input generation launches hundreds of torch kernels first)
#   prof_TAG_{net,hot,scan}.ncu-rep   one `--set full` capture of the first launch of each kernel (full 64 800-cell grid)
# Numbers printed under ncu are never bench values.
TAG=$1
CMD="python bench.py --no-e2e --no-cpu --no-extra --steps 1 --warmup 1"
$CMD > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|hdp" -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
for k in net:k_thr_net hot:k_hot_words scan:k_scan; do
  short=${k%%:*}; name=${k##*:}
  # (sources are imported for the threshold kernel only: with them the three reports exceed what gpurun carries back)
  SRC=""; [ "$short" = "net" ] && SRC="--import-source on"
  ncu --set full --clock-control none $SRC -k regex:"$name" -c 1 -f -o gpurun_out/prof_${TAG}_$short $CMD > gpurun_out/ncu_${TAG}_$short.log 2>&1
  # every report embeds the library's whole cubin (~30 MB): keep its pages as CSV and drop it, or the three exceed what gpurun carries back
  ncu -i gpurun_out/prof_${TAG}_$short.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}_$short.csv 2>/dev/null
  [ "$short" = "net" ] && ncu -i gpurun_out/prof_${TAG}_$short.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/src_${TAG}_$short.csv 2>/dev/null
  rm -f gpurun_out/prof_${TAG}_$short.ncu-rep
done
ls -la gpurun_out/*$TAG*
