#!/bin/bash
timeout 600 python scripts/debug_localdev.py 2>&1 | tail -25
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8
