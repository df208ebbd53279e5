#!/usr/bin/env bash
# Build the UNMODIFIED NebulaSEM reference (read in place from $NSEM_REFERENCE, default
# /root/reference) into oracle/_ref/ as the parity oracle and CPU baseline.
# TEST INFRASTRUCTURE ONLY: nothing in nebulasem_b200/ links or calls these binaries.
#
#   oracle/_ref/parity/{euler,convection,mesh,prepare,geomdump,refinedump}   -O2 -ffp-contract=off          (parity oracle)
#   oracle/_ref/fast/{euler,mesh}                       -O3 -funroll-loops -march=x86-64-v3 -fopenmp
#                                                        (the reference's release flags, CMakeLists.txt:36-37,
#                                                         with a portable -march so the binary also runs on the GPU box)
# The reference's own build system (cmake + find_package(MPI)) is NOT used: the image has no MPI,
# so the sources are compiled directly with g++ against oracle/mpi_shim/mpi.h (single rank).
# No reference source is copied into the repository; only objects/binaries land in oracle/_ref/ (git-ignored).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
R="${NSEM_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
JOBS="${JOBS:-$(nproc)}"
if [ ! -d "$R/src/field" ]; then
    echo "build_ref: reference tree $R not present; keeping prebuilt $OUT" >&2
    exit 0
fi
INC="-I$HERE/mpi_shim -I$R/src/field -I$R/src/mesh -I$R/src/mp -I$R/src/prepare -I$R/src/solvers -I$R/src/tensor \
 -I$R/src/turbulence -I$R/src/turbulence/ke -I$R/src/turbulence/kw -I$R/src/turbulence/les \
 -I$R/src/turbulence/mixing_length -I$R/src/turbulence/realizableke -I$R/src/turbulence/rngke \
 -I$R/src/util -I$R/src/vtk -I$R/apps/utils"
LIBSRC="$R/src/field/dg.cpp $R/src/field/field.cpp $R/src/mesh/hexMesh.cpp $R/src/mesh/mesh.cpp $R/src/mesh/mshMesh.cpp
 $R/src/mp/mp.cpp $R/src/prepare/prepare.cpp $R/src/solvers/solve.cpp $R/src/tensor/tensor.cpp $R/src/util/util.cpp
 $R/src/vtk/vtk.cpp $R/src/turbulence/turbulence.cpp $(ls $R/src/turbulence/*/*.cpp) $(ls $R/apps/utils/*.cpp)"

build_variant() {
    local name="$1"; shift
    local flags="$*"
    local dir="$OUT/$name"
    local stamp="$dir/.flags"
    if [ -f "$stamp" ] && [ "$(cat "$stamp")" = "$flags" ] && [ -x "$dir/euler" ] && [ -x "$dir/mesh" ] \
       && [ "$HERE/tools/geomdump.cpp" -ot "$dir/euler" ] && [ -x "$dir/refinedump" ] && [ "$HERE/tools/refinedump.cpp" -ot "$dir/refinedump" ] && [ -x "$dir/convection" ]; then
        echo "build_ref: $name up to date"; return
    fi
    echo "build_ref: compiling $name ($flags)"
    mkdir -p "$dir/obj"
    local pids=() n=0
    for s in $LIBSRC; do
        local o="$dir/obj/$(echo "$s" | sed "s#$R/##; s#/#_#g; s#\.cpp\$#.o#")"
        g++ -std=c++17 $flags -DUSE_DOUBLE -DUSE_EXPR_TMPL -w $INC -c "$s" -o "$o" &
        pids+=($!); n=$((n+1))
        if [ $n -ge $JOBS ]; then wait "${pids[0]}"; pids=("${pids[@]:1}"); n=$((n-1)); fi
    done
    wait
    ar rcs "$dir/libnebulasem.a" "$dir"/obj/*.o
    g++ -std=c++17 $flags -DUSE_DOUBLE -DUSE_EXPR_TMPL -w $INC "$R/apps/euler/euler.cpp" "$dir/libnebulasem.a" -o "$dir/euler" &
    g++ -std=c++17 $flags -DUSE_DOUBLE -DUSE_EXPR_TMPL -w $INC "$R/apps/mesh/meshApp.cpp" "$dir/libnebulasem.a" -o "$dir/mesh" &
    g++ -std=c++17 $flags -DUSE_DOUBLE -DUSE_EXPR_TMPL -w $INC "$R/apps/convection/convection.cpp" "$dir/libnebulasem.a" -o "$dir/convection" &
    g++ -std=c++17 $flags -DUSE_DOUBLE -DUSE_EXPR_TMPL -w $INC "$R/apps/prepare/prepareApp.cpp" "$dir/libnebulasem.a" -o "$dir/prepare" &
    g++ -std=c++17 $flags -DUSE_DOUBLE -DUSE_EXPR_TMPL -w $INC "$HERE/tools/geomdump.cpp" "$dir/libnebulasem.a" -o "$dir/geomdump" &
    g++ -std=c++17 $flags -DUSE_DOUBLE -DUSE_EXPR_TMPL -w $INC "$HERE/tools/refinedump.cpp" "$dir/libnebulasem.a" -o "$dir/refinedump" &
    wait
    rm -rf "$dir/obj" "$dir/libnebulasem.a"      # keep oracle/_ref small: it travels to the GPU box with every gpurun
    echo "$flags" > "$stamp"
}

build_variant parity "-O2 -ffp-contract=off"
build_variant fast "-O3 -funroll-loops -march=x86-64-v3 -fopenmp"
ls -la "$OUT"/parity "$OUT"/fast | grep -v '\.o$'
