#!/bin/bash
# usage: build.sh name -> /root/repo/ab_libs/name.so ; prints spills of ring kernel 1280
name=$1; d=/tmp/var/$name; mkdir -p $d/obj /root/repo/ab_libs
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v -diag-suppress=20013,20015 -I $d/include"
pids=()
for s in engine decode mel enc_gemm enc_attn enc_gemm_tc; do
  nvcc $FLAGS -c $d/csrc/$s.cu -o $d/obj/$s.o > $d/obj/$s.log 2>&1 & pids+=($!)
done
fail=0; for p in "${pids[@]}"; do wait $p || fail=1; done
if [ $fail = 1 ]; then grep -h "error" $d/obj/*.log | head; echo "[$name] BUILD FAILED"; exit 1; fi
nvcc -shared -o /root/repo/ab_libs/$name.so $d/obj/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
grep -A2 "Function properties for _ZN2wm25dec_iteration_ring_kernelILi1280ELb0" $d/obj/decode.log | grep -E "spill|registers" | tr '\n' ' ' | sed "s/^/[$name] /"; echo
python /tmp/codesize.py $d/obj/decode.o | head -1 | sed "s/^/[$name] /"
