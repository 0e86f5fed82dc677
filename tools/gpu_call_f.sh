#!/bin/bash
bash tools/run_tp.sh 4 7b --steps 64 --warmup 8 --no-per-op
bash tools/run_tp.sh 4 33b_g32_act --model 33b --groupsize 32 --act-order --steps 16 --warmup 4 --no-per-op
