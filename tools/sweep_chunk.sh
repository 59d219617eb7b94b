for n in 4 16 64 512; do
  HDP_B200_THR_CHUNK_GROUPS=$n python bench.py --no-e2e --no-cpu --steps 3 2>/dev/null > gpurun_out/sw_$n.json
  python -c "
import json; d=json.load(open('gpurun_out/sw_$n.json')); print($n, d['ms_per_step'], d['roofline']['kernel_ms'])"
done
