#!/bin/bash
# build_variant.sh NAME "EXTRA nvcc flags": builds csrc into /root/repo/variants/NAME/libmicrosound_b200.so (A/B tests)
set -e
name=$1; extra=$2
src=/root/repo/audio_suite_b200/csrc
out=/root/repo/variants/$name
mkdir -p $out/csrc $out/include
cp $src/*.cu $src/*.cuh $src/*.h $src/*.inl $src/Makefile $out/csrc/
cp /root/repo/include/*.h $out/include/
sed -i 's#../../include/microsound_b200.h#../include/microsound_b200.h#' $out/csrc/ms_prelude.h
sed -i 's#../../include/\*.h#../include/*.h#' $out/csrc/Makefile
make -s -j4 -C $out/csrc libmicrosound_b200.so EXTRA="$extra"
cp $out/csrc/libmicrosound_b200.so $out/
rm -rf $out/csrc $out/include
ls -la $out
