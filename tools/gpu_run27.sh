#!/bin/bash
for na in 2 4 8; do echo "--- NO_EPI3 NACC=$na"; VCG_NO_EPI3=1 VCG_NACC=$na timeout 300 python tools/bench_conv.py e0_b64 d5_b64 dU4_b64 dU3_b64 2>&1 | tail -4 | cut -c1-150; done
