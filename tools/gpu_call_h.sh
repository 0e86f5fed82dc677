#!/bin/bash
bash tools/run_tp.sh 8 65b_seq4096 --model 65b --seq 4096 --steps 16 --warmup 4 --no-per-op --no-prefill
