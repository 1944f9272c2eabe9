#!/bin/bash
mkdir -p gpurun_out
UOC_CONV_LAYERS_ONLY="l3 3x3 256 d2,l4 1x1 down" UOC_CONV_LAYERS_VARIANTS="UOC_CONV_TRACE=1,UOC_CONV_PAIR=1" timeout 300 python tools/conv_layers.py 2 8 2>&1 | grep "conv trace" | awk 'NR%7==1' > gpurun_out/r2s9_conv_trace.txt
cat gpurun_out/r2s9_conv_trace.txt
