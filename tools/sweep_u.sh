#!/bin/bash
# developer sweep: gather batch 4 (default build) vs 8 (build/libamgcl_b200_u8.so)
for v in u4 u8; do
  if [ $v = u8 ]; then export B200_LIB=$PWD/build/libamgcl_b200_u8.so; fi
  echo "== $v"
  timeout 500 python tools/gpu_check.py spmv 256 2>&1 | grep "^{" > gpurun_out/spmv256_$v.log
  python - <<PY
import json
for l in open("gpurun_out/spmv256_$v.log"):
    d=json.loads(l)
    if "variant" in d: print(d["variant"],d["nnz_cap"],d["ctas_per_sm"],d["stages"],d["spmv"]["GBs"],d["residual"]["GBs"],d["relax"]["GBs"])
    else: print(d)
PY
  timeout 600 python tools/gpu_check.py levels 192 > gpurun_out/levels192_$v.log 2>&1
  grep BEST gpurun_out/levels192_$v.log | cut -c1-220
done
