#!/bin/bash
mkdir -p gpurun_out
echo "== default"; timeout 300 python tools/smoke_diag.py 2>&1 | grep -v Warning | tail -20
for e in "VCG_TC2=0" "VCG_NO_EPI3=1" "VCG_NO_TSTORE=1" "VCG_WTC2=0" "VCG_NO_RING=1"; do
  echo "== $e"; env $e SMOKE_ORACLE=0 timeout 300 python tools/smoke_diag.py 2>&1 | tail -1
done
echo "== default again"; SMOKE_ORACLE=0 timeout 300 python tools/smoke_diag.py 2>&1 | tail -1
