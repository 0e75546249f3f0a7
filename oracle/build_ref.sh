#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- builds oracle/_ref/stcsp_ref: the reference's own solver, compiled from the
# sources where they lie under /root/reference/src (nothing is copied into the repo), linked with the
# hand-written stand-in front end oracle/ref_frontend.cpp (lex/yacc are absent in this image) and the
# hand-written token header oracle/refshim/y.tab.h.  Output goes only to oracle/_ref/ (git-ignored,
# but NOT gpurun-ignored: the binary travels to the GPU box, where /root/reference does not exist).
#
# -std=gnu++98 : the reference uses __gnu_cxx::hash_map with `using namespace` of both std and __gnu_cxx
#                (src/graph.h:8-9,31); `hash<int>` is ambiguous from C++11 on.
# --wrap=malloc: zero-fill shim, see ref_frontend.cpp.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="${STCSP_REFERENCE_SRC:-/root/reference/src}"
out="$here/_ref"
if [ ! -d "$ref" ]; then
    echo "build_ref: $ref not present (GPU box?) - keeping prebuilt $out/stcsp_ref" >&2
    [ -x "$out/stcsp_ref" ] && exit 0 || exit 1
fi
mkdir -p "$out"
srcs=(constraint graph node solver solveralgorithm token util variable)
objs=()
for s in "${srcs[@]}"; do
    g++ -std=gnu++98 -O2 -w -I"$here/refshim" -I"$ref" -c "$ref/$s.cpp" -o "$out/$s.o"
    objs+=("$out/$s.o")
done
g++ -std=gnu++98 -O2 -w -I"$here/refshim" -I"$ref" -c "$here/ref_frontend.cpp" -o "$out/ref_frontend.o"
g++ -O2 "${objs[@]}" "$out/ref_frontend.o" -Wl,--wrap=malloc -o "$out/stcsp_ref"
echo "built $out/stcsp_ref"
# stcsp_ref_gpu: the same reference objects (front end stand-in, normaliser, post-processing, DOT writer) with the ONE call
# solverSolve(Solver*, bool) redirected by the linker to oracle/gpu_bridge.cpp, i.e. to stcsp_gpu_solve of the product
# library (the compiled form of the binding shown in INTEGRATION.md).  Needs the product library to be built first.
lib="$here/../stcsp_solver_b200"
if [ -f "$lib/libstcsp_b200.so" ]; then
    g++ -std=gnu++98 -O2 -w -I"$here/refshim" -I"$ref" -I"$here/../include" -c "$here/gpu_bridge.cpp" -o "$out/gpu_bridge.o"
    g++ -O2 "${objs[@]}" "$out/ref_frontend.o" "$out/gpu_bridge.o" -Wl,--wrap=malloc -Wl,--wrap=_Z11solverSolveP6Solverb \
        -L"$lib" -lstcsp_b200 -Wl,-rpath,'$ORIGIN/../../stcsp_solver_b200' -o "$out/stcsp_ref_gpu"
    echo "built $out/stcsp_ref_gpu"
else
    echo "build_ref: $lib/libstcsp_b200.so not built yet - skipping stcsp_ref_gpu" >&2
fi
# stcsp_ref_bridge_cpu: the same bridge with the CPU oracle behind the three C-ABI symbols (bridge_cpu_shim.cpp), so the
# bridge code itself is covered by the CPU test suite
if [ -f "$here/liboracle.so" ]; then
    g++ -std=gnu++98 -O2 -w -I"$here/refshim" -I"$ref" -I"$here/../include" -c "$here/gpu_bridge.cpp" -o "$out/gpu_bridge.o"
    g++ -O2 -I"$here/../include" -c "$here/bridge_cpu_shim.cpp" -o "$out/bridge_cpu_shim.o"
    g++ -O2 "${objs[@]}" "$out/ref_frontend.o" "$out/gpu_bridge.o" "$out/bridge_cpu_shim.o" -Wl,--wrap=malloc \
        -Wl,--wrap=_Z11solverSolveP6Solverb -L"$here" -loracle -Wl,-rpath,'$ORIGIN/..' -o "$out/stcsp_ref_bridge_cpu"
    echo "built $out/stcsp_ref_bridge_cpu"
fi
rm -f "$out"/*.o
