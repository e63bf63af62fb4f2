#!/bin/bash
# usage: profiles/build_exp.sh <name> "<extra nvcc flags>"  -> profiles/r02/libgnssacq_<name>.so (experiment build, not product)
set -e
N=$1; F=$2
make -C /root/repo/assignment-for-aae6102_gnss-sdr_b200/csrc -j4 BUILD=/tmp/exp_$N OUT=/root/repo/profiles/r02/libgnssacq_$N.so EXTRA="$F" > /tmp/exp_$N.log 2>&1 || (tail -20 /tmp/exp_$N.log; exit 1)
ls -la /root/repo/profiles/r02/libgnssacq_$N.so
