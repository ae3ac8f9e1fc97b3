#!/bin/bash
timeout 600 python scripts/debug_localdev.py > gpurun_out/localdev.log 2>&1; tail -3 gpurun_out/localdev.log
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6
