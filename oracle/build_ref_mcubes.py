#!/usr/bin/env python
"""Test infrastructure: compile the REFERENCE's marching cubes core (libmcubes/marchingcubes.{h,cpp}, the code behind
libmcubes.marching_cubes used by reconstruct.py:190) from where it lies under /root/reference into oracle/_ref/
(git-ignored), behind a small C shim written here (the reference's own Python wrapper does not build against numpy 2).
Only the two reference files are compiled; no reference source is copied into the repository.

    python oracle/build_ref_mcubes.py      # -> oracle/_ref/libref_mcubes.so
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = "/root/reference/reg_slices/src_convonet/utils/libmcubes"

SHIM = r'''
// C shim over the reference's mc::marching_cubes<double> exactly as pywrapper.cpp:90-128 calls it
// (lower = 0, upper = shape - 1, functor taking INT coordinates, i.e. the doubles are truncated).
#include <cstdlib>
#include <cstring>
#include <vector>
#include "marchingcubes.h"
struct Vol {
  const double* v; long ny, nz;
  double operator()(int x, int y, int z) const { return v[((long)x * ny + y) * nz + z]; }
};
extern "C" {
int ref_mc_run(const double* vol, long nx, long ny, long nz, double iso, double** verts, size_t* nverts, size_t** tris,
               size_t* ntris) {
  double lower[3] = {0, 0, 0};
  double upper[3] = {(double)(nx - 1), (double)(ny - 1), (double)(nz - 1)};
  std::vector<double> vertices;
  std::vector<size_t> polygons;
  mc::marching_cubes<double>(lower, upper, (int)nx, (int)ny, (int)nz, Vol{vol, ny, nz}, iso, vertices, polygons);
  *nverts = vertices.size();
  *ntris = polygons.size();
  *verts = (double*)std::malloc(sizeof(double) * (vertices.size() + 1));
  *tris = (size_t*)std::malloc(sizeof(size_t) * (polygons.size() + 1));
  if (!vertices.empty()) std::memcpy(*verts, vertices.data(), sizeof(double) * vertices.size());
  if (!polygons.empty()) std::memcpy(*tris, polygons.data(), sizeof(size_t) * polygons.size());
  return 0;
}
void ref_mc_free(void* p) { std::free(p); }
}
'''


def build():
    if not os.path.exists(os.path.join(REF, "marchingcubes.cpp")):
        return None
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "libref_mcubes.so")
    if os.path.exists(so):
        return so
    shim = os.path.join(OUT, "mc_shim.cpp")
    with open(shim, "w") as f:
        f.write(SHIM)
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-w", "-I", REF, os.path.join(REF, "marchingcubes.cpp"), shim,
                           "-o", so])
    os.remove(shim)
    return so


def run(vol, iso):
    """numpy float64 (nx,ny,nz) -> (vertices (n,3) float64, triangles (m,3) uint64) from the reference."""
    import ctypes as C
    import numpy as np
    L = C.CDLL(build())
    vol = np.ascontiguousarray(vol, dtype=np.float64)
    pv, pt = C.POINTER(C.c_double)(), C.POINTER(C.c_size_t)()
    nv, nt = C.c_size_t(), C.c_size_t()
    L.ref_mc_run(vol.ctypes.data_as(C.POINTER(C.c_double)), C.c_long(vol.shape[0]), C.c_long(vol.shape[1]),
                 C.c_long(vol.shape[2]), C.c_double(iso), C.byref(pv), C.byref(nv), C.byref(pt), C.byref(nt))
    v = np.ctypeslib.as_array(pv, shape=(max(nv.value, 1),))[:nv.value].copy().reshape(-1, 3)
    t = np.ctypeslib.as_array(pt, shape=(max(nt.value, 1),))[:nt.value].copy().reshape(-1, 3)
    L.ref_mc_free(pv)
    L.ref_mc_free(pt)
    return v, t


if __name__ == "__main__":
    print(build())
