#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.
# Compiles the reference's own CPU path tracer (32 translation units, where they lie under
# $REF, default /root/reference) plus oracle/ref_driver.cpp into oracle/_ref/ref_driver.
# Outputs go ONLY into oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).
# Recipe proven in SURVEY.md section 8c.  The reference's own build system is not used.
#
# Two sed edits are applied to a COPY of src/bvh.cpp (oracle/_ref/patched/bvh.cpp):
#   1. clamp the SAH bucket index (bvh.cpp:47 overflows the 32-entry bucket vector when a
#      centroid sits on the node's upper bound -> glibc abort on every mesh scene);
#   2. two counters in the BVHAccel::intersect overloads (bvh.cpp:331-363) so the
#      reference's segment throughput can be reported.
# Neither changes closest-hit results.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "[build_ref] $REF not present; keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj" "$OUT/patched" "$OUT/scenes"

CXX="${CXX:-g++}"
FLAGS="-std=gnu++11 -O2 -w -fpermissive -include cstdint -include cstring -DGLEW_NO_GLU \
 -I$REF/CMU462/include -I$REF/CMU462/include/CMU462 -I$REF/src"

sed -e 's|int bIndex = (c - lb) / interval;|int bIndex = (c - lb) / interval; if (bIndex > bucketNum-1) bIndex = bucketNum-1; if (bIndex < 0) bIndex = 0;|' \
    -e '1s|^|extern long long g_ref_any_calls, g_ref_closest_calls;\n|' \
    -e 's|return node_intersect(get_root(), primitives, ray);|::g_ref_any_calls++; return node_intersect(get_root(), primitives, ray);|' \
    -e 's|return node_intersect(get_root(), primitives, ray, i);|::g_ref_closest_calls++; return node_intersect(get_root(), primitives, ray, i);|' \
    "$REF/src/bvh.cpp" > "$OUT/patched/bvh.cpp"
grep -q "bIndex > bucketNum-1" "$OUT/patched/bvh.cpp"
grep -q "g_ref_any_calls" "$OUT/patched/bvh.cpp"
grep -q "g_ref_closest_calls" "$OUT/patched/bvh.cpp"

TUS="CMU462/src/vector2D.cpp CMU462/src/vector3D.cpp CMU462/src/vector4D.cpp CMU462/src/matrix3x3.cpp
 CMU462/src/matrix4x4.cpp CMU462/src/spectrum.cpp CMU462/src/color.cpp CMU462/src/tinyxml2.cpp CMU462/src/lodepng.cpp
 src/collada/collada.cpp src/collada/camera_info.cpp src/collada/light_info.cpp src/collada/sphere_info.cpp
 src/collada/polymesh_info.cpp src/collada/material_info.cpp
 src/dynamic_scene/mesh.cpp src/dynamic_scene/scene.cpp src/dynamic_scene/sphere.cpp
 src/static_scene/sphere.cpp src/static_scene/triangle.cpp src/static_scene/object.cpp
 src/static_scene/environment_light.cpp src/static_scene/light.cpp
 src/halfEdgeMesh.cpp src/meshEdit.cpp src/bbox.cpp src/bsdf.cpp src/camera.cpp src/sampler.cpp
 src/pathtracer.cpp src/misc/sphere_drawing.cpp"

pids=()
for tu in $TUS; do
  o="$OUT/obj/$(echo "$tu" | tr '/' '_' | sed 's/\.cpp$/.o/')"
  if [ ! -f "$o" ] || [ "$REF/$tu" -nt "$o" ]; then
    $CXX $FLAGS -c "$REF/$tu" -o "$o" &
    pids+=($!)
  fi
done
$CXX $FLAGS -c "$OUT/patched/bvh.cpp" -o "$OUT/obj/src_bvh_patched.o" &
pids+=($!)
$CXX $FLAGS -fno-access-control -c "$HERE/ref_driver.cpp" -o "$OUT/obj/ref_driver.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done

$CXX "$OUT"/obj/*.o -o "$OUT/ref_driver" -lpthread -Wl,--unresolved-symbols=ignore-all

# The reference-side binding, compiled against the reference's real PathTracer: integration/cuda_path_tracer_shim.{h,cpp}
# (drop-in for cuda_src/setup.{h,cu}) + integration/ref_gpu_driver.cpp (Application::startGPURayTracing restated) + the same
# reference objects, linked against the product's libdsrt.so.  -fno-access-control only for EnvironmentLight::envMap
# (environment_light.h:50, private); everything else the shim reads is public in the reference's headers.
REPO="$(cd "$HERE/.." && pwd)"
if [ -f "$REPO/dsgpuraytracing_b200/libdsrt.so" ]; then
  mkdir -p "$OUT/obj_gpu"
  $CXX $FLAGS -fno-access-control -I"$REPO/include" -c "$REPO/integration/cuda_path_tracer_shim.cpp" -o "$OUT/obj_gpu/shim.o"
  $CXX $FLAGS -fno-access-control -I"$REPO/include" -c "$REPO/integration/ref_gpu_driver.cpp" -o "$OUT/obj_gpu/ref_gpu_driver.o"
  REFOBJS=$(ls "$OUT"/obj/*.o | grep -v "ref_driver.o")
  $CXX "$OUT/obj_gpu/shim.o" "$OUT/obj_gpu/ref_gpu_driver.o" $REFOBJS -o "$OUT/ref_gpu_driver" -lpthread \
       -L"$REPO/dsgpuraytracing_b200" -ldsrt -Wl,-rpath,'$ORIGIN/../../dsgpuraytracing_b200' -Wl,--unresolved-symbols=ignore-all
  echo "[build_ref] built $OUT/ref_gpu_driver (reference PathTracer -> CUDAPathTracer shim -> libdsrt.so)"
fi
# the reference's vendored tinyexr behind a tiny command-line tool: pins the product's own OpenEXR reader
$CXX -std=gnu++11 -O2 -w -fpermissive -I$REF/CMU462/include/CMU462 "$HERE/exr_ref.cpp" -o "$OUT/exr_ref"

# stage the scene/camera DATA files next to the binary so GPU-box tests can use them
# (data, not source; still kept out of git history by .gitignore)
for f in CBspheres.dae CBspheres_lambertian.dae CBempty.dae CBgems.dae CBcoil.dae CBbunny.dae bunny.dae; do
  cp -f "$REF/dae/sky/$f" "$OUT/scenes/$f"
done
cp -f "$REF/cam_dragon.info" "$REF/cam_sphere.info" "$OUT/scenes/"
echo "[build_ref] built $OUT/ref_driver"
