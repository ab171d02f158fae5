#!/bin/bash
# BF16 error of the committed benchmark cases and of ResNet variants (tools/bf16_sweep.py), then the conv parity tests
mkdir -p gpurun_out
timeout 900 python tools/bf16_sweep.py --cases ${SWEEP:-resnet50:0.1:76 resnet50:0.25:64 resnet50:0.25:48 resnet50:1.0:48 resnet18:1.0:48} 2>&1 | grep -E "^bench|^resnet|Error|error" | tail -20
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu -x 2>&1 | tail -5
