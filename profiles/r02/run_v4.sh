#!/bin/bash
O=gpurun_out/r02v6; mkdir -p $O
for v in mb9 mb10 mb11; do
  lib=$PWD/_variants/librtb200_$v.so; [ $v = default ] && lib=$PWD/uu-infogr-raytracer_b200/librtb200.so
  for s in config3 config4; do
    echo "$v $s: $(RTB200_LIB=$lib timeout 120 python profiles/prof_driver.py 1 6 $s 2>&1 | tail -2 | tr '\n' ' ')"
  done
done | tee $O/variants.txt
