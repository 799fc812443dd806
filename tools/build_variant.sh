#!/bin/bash
# usage: tools/build_variant.sh <name> <nvcc -D flags...>   -> libgoldilocks_b200/_exp_<name>.so (kernel experiments; select with GOLDILOCKS_B200_LIB)
name=$1; shift
make -j16 lib OBJDIR=build/exp_$name LIB=libgoldilocks_b200/_exp_$name.so EXTRA_NVCCFLAGS="$*" 2>&1 | grep -E "error" -A3
ls -la libgoldilocks_b200/_exp_$name.so
