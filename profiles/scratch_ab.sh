# A/B timing of library variants / env switches: bench.py kernel_ms per variant (scratch tool)
# usage: scratch_ab.sh "label ENV=1 ENV2=1" ...     (TBIRCH_LIB=v_<name> picks ternary_birch_b200/libtb_v_<name>.so)
for spec in "$@"; do
  set -- $spec; label=$1; shift
  envs=""
  for e in "$@"; do case $e in LIB=*) envs="$envs TBIRCH_LIB=$PWD/ternary_birch_b200/libtb_v_${e#LIB=}.so";; *) envs="$envs $e";; esac; done
  env A=1 $envs python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label', 'kernel_ms', round(r['roofline']['kernel_ms'],3), 'rows', round(r['roofline']['rows_kernels_ms'],3), 'value', round(r['value']/1e9,3))"
done
