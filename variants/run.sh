for v in variants/libA.so variants/libB.so variants/libC.so variants/libD.so; do
  echo "== variant: ${v:-default}"
  TA_EVAL_LIB=${v:+$PWD/$v} python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step']); print({k:round(v['ms'],3) for k,v in d['roofline']['stages'].items()})"
done
