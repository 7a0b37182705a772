#!/bin/bash
# round 2, session 2: scalar LayerNorm kernel writes F16F8 rows in the interleaved layout (was still the separate-plane one, and its
# zero padding was the F16X2 one) -- the new test first, then pass Z (full GPU suite, smoke, bench with the driver's flags)
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gemm_f8.py -q -m gpu -x -k "scalar_path" 2>&1 | tail -2
sed -e 's/r03z_/r03z5_/g' scripts/gpu_r03_z.sh | grep -v "impl reference" > /tmp/z5.sh; bash /tmp/z5.sh
