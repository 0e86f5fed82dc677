#!/bin/bash
ORACLE=1 timeout 400 python tools/tp_debug.py --model 33b --groupsize 32 --act-order --layers 2 2>&1 | grep -v Warn | tail -16
