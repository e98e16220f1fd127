# timing of the v2 fused kernel with parts switched off (results invalid; diagnostic only)
for d in "$@"; do
  echo "== SNN_B200_DEBUG=$d"
  SNN_B200_DEBUG=$d SNN_B200_PROF=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | tail -50 | cut -c1-170 | grep -E "exchange|gather|early|barrier|late pass|S1|per step|slowest group|step 1"
done
