#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu 2>&1 | tail -6
