#!/bin/bash
bash tools/run_tp.sh 2 7b --steps 64 --warmup 8 --no-per-op --no-prefill
