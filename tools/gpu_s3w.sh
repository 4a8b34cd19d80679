#!/bin/bash
# session 3, call W: mixed-size request vs uniform request
timeout 100 python tools/mixed_bench.py 2>&1 | grep -E "uniform|mixed"
