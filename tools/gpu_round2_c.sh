#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "layer_major or golden or oracle" 2>&1 | tail -5 > gpurun_out/c_lbwd_test.log
cat gpurun_out/c_lbwd_test.log
python tools/ab_lbwd.py 2>&1 | grep -E "rep 0|rep 5|rror" 
bash tools/gpu_variants.sh 2>&1 | grep -E "^==|layer-major|eager|rror"
python tools/trace_lbwd.py > gpurun_out/trace_lbwd.log 2>&1
