#!/bin/bash
# Builds ab_libs/libern_tracewaits.so: the in-tree library with -DERN_SIM_TRACE_WAITS in the scoring kernel
# (clock reads around every wait of the MMA thread and epilogue warp 0; ~8 % slower; for tools/trace_sim.py only).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
CSRC="$ROOT/fashionern_aaai2024_b200/csrc"
make -C "$CSRC" -j8 > /dev/null
mkdir -p "$ROOT/ab_libs"
F="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -cudart static"
nvcc $F -DERN_SIM_TRACE_WAITS -c "$CSRC/ern_sim_tc.cu" -o /tmp/ern_sim_tc_trace.o
others=$(ls "$CSRC"/build/*.o | grep -v ern_sim_tc.o)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart static -o "$ROOT/ab_libs/libern_tracewaits.so" $others /tmp/ern_sim_tc_trace.o
echo "built $ROOT/ab_libs/libern_tracewaits.so"
