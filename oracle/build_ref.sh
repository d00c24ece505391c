#!/usr/bin/env bash
# oracle/build_ref.sh -- build the UNMODIFIED reference `bcalm` binary from /root/reference into oracle/_ref/.
#
# The reference cannot be compiled from "a few source files": gatb-core needs its vendored HDF5 (configured and built by
# cmake, with generated headers).  We therefore drive the reference's own CMakeLists out-of-tree in a scratch directory
# (SURVEY.md 8c established that this works in this image: cmake 4.x needs the policy flag) and copy ONLY the resulting
# executable into oracle/_ref/.  No reference source is copied into the repo.  If /root/reference or cmake is missing
# (e.g. on the GPU box) this script is a no-op: the prebuilt oracle/_ref/bcalm travels with the snapshot.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="${CDBG_REFERENCE_DIR:-/root/reference}"
out="$here/_ref"
if [ -x "$out/bcalm" ]; then echo "oracle/_ref/bcalm already present"; exit 0; fi
if [ ! -f "$ref/CMakeLists.txt" ] || ! command -v cmake >/dev/null; then
  echo "reference sources or cmake not available: skipping (oracle/_ref stays as is)"; exit 0
fi
build="${CDBG_REF_BUILD_DIR:-/tmp/cdbg_refbuild}"
mkdir -p "$build" "$out"
cd "$build"
cmake "$ref" -DCMAKE_POLICY_VERSION_MINIMUM=3.5 -DCMAKE_BUILD_TYPE=Release > cmake.log 2>&1
make -j"$(nproc)" bcalm > make.log 2>&1
cp "$build/bcalm" "$out/bcalm"
echo "built $out/bcalm"
