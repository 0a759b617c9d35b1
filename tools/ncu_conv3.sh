#!/bin/bash
# ncu capture of one k x k conv layer through the engine (halo mode on/off via env): tools/ncu_conv3.sh <tag> [env...]
tag=$1; shift
env "$@" python tools/bench_conv.py 96 24 3 8 240 240 > gpurun_out/${tag}_plain.log 2>&1 &&
env "$@" ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 4 -c 1 -o gpurun_out/${tag} -f python tools/bench_conv.py 96 24 3 8 240 240 > gpurun_out/${tag}_ncu.log 2>&1
