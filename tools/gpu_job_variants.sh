#!/bin/bash
# A/B of experiment builds (libh263cu_<name>.so, see build.build_variant): bench only.
# usage: tools/gpu_job_variants.sh name1 name2 ...   ("default" = the product library)
set +e
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = "default" ]; then unset H263CU_LIB; else export H263CU_LIB=$PWD/h263_rs_b200/libh263cu_$v.so; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --skip-extras > gpurun_out/variant_$v.json 2> gpurun_out/variant_$v.err
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open('gpurun_out/variant_%s.json'%v))
    print("%-12s kernel_ms %.4f  frac %.3f  step_ms %.4f"%(v,d['roofline']['kernel_ms'],d['roofline']['frac'],d['ms_per_step']))
except Exception as e:
    print(v,'FAILED',e); print(open('gpurun_out/variant_%s.err'%v).read()[-500:])
PY
done
