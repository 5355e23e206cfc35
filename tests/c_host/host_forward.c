/* A C host of libvda's handle-level API (include/vda.h: vda_create ... vda_forward): what a non-Python caller of
 * VideoDepthAnything.forward (video_depth_anything/video_depth.py:89-164) links against.  Test infrastructure:
 * tests/test_c_host.py builds it with gcc, feeds it a state dict and an input window through plain binary files and
 * requires the depth map it writes to be bit-identical to the Python engine's.
 *
 *   host_forward <encoder> <features> <oc0> <oc1> <oc2> <oc3> <dtype> <weights.bin> <input.bin> <output.bin>
 *
 * weights.bin: int32 count, then per tensor: int32 name_len, name bytes, int32 ndim, int64 shape[ndim], float data[]
 * input.bin:   int32 B, T, H, W, then float x[B*T*3*H*W]          output.bin: float depth[B*T*H*W]
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vda.h"

#define DIE(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return 1; } while (0)
#define VDA(call) do { if ((call) != 0) DIE("%s failed: %s", #call, vda_last_error()); } while (0)
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) DIE("%s: %s", #call, cudaGetErrorString(e_)); } while (0)

int main(int argc, char** argv) {
  if (argc == 2 && strcmp(argv[1], "--version") == 0) {   /* link check without a GPU */
    printf("%d\n", vda_version());
    return 0;
  }
  if (argc != 11) DIE("usage: %s encoder features oc0 oc1 oc2 oc3 dtype weights.bin input.bin output.bin", argv[0]);
  const int32_t oc[4] = {atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6])};
  vda_model* m = NULL;
  VDA(vda_create(argv[1], atoi(argv[2]), oc, 32, atoi(argv[7]), 0, &m));

  FILE* f = fopen(argv[8], "rb");
  if (!f) DIE("cannot open %s", argv[8]);
  int32_t count = 0;
  if (fread(&count, 4, 1, f) != 1) DIE("bad weights file");
  for (int i = 0; i < count; ++i) {
    int32_t nl = 0, nd = 0;
    char name[512];
    int64_t shape[8];
    if (fread(&nl, 4, 1, f) != 1 || nl <= 0 || nl >= 512 || fread(name, 1, (size_t)nl, f) != (size_t)nl) DIE("bad tensor name");
    name[nl] = 0;
    if (fread(&nd, 4, 1, f) != 1 || nd < 0 || nd > 8 || fread(shape, 8, (size_t)nd, f) != (size_t)nd) DIE("bad tensor shape");
    size_t n = 1;
    for (int d = 0; d < nd; ++d) n *= (size_t)shape[d];
    float* data = (float*)malloc(n * 4 + 4);
    if (!data || fread(data, 4, n, f) != n) DIE("bad tensor data (%s)", name);
    VDA(vda_set_weight(m, name, data, shape, nd));
    free(data);
  }
  fclose(f);
  VDA(vda_finalize_weights(m));

  f = fopen(argv[9], "rb");
  if (!f) DIE("cannot open %s", argv[9]);
  int32_t dims[4];
  if (fread(dims, 4, 4, f) != 4) DIE("bad input file");
  const int B = dims[0], T = dims[1], H = dims[2], W = dims[3];
  const size_t nin = (size_t)B * T * 3 * H * W, nout = (size_t)B * T * H * W;
  float* hx = (float*)malloc(nin * 4);
  if (!hx || fread(hx, 4, nin, f) != nin) DIE("bad input data");
  fclose(f);

  float *dx = NULL, *dd = NULL;
  void* ws = NULL;
  const int64_t ws_bytes = vda_workspace_bytes(m, B, T, H, W);
  if (ws_bytes < 0) DIE("vda_workspace_bytes: %s", vda_last_error());
  CU(cudaMalloc((void**)&dx, nin * 4));
  CU(cudaMalloc((void**)&dd, nout * 4));
  CU(cudaMalloc(&ws, (size_t)ws_bytes));
  CU(cudaMemcpy(dx, hx, nin * 4, cudaMemcpyHostToDevice));
  cudaStream_t st;
  CU(cudaStreamCreate(&st));
  /* a too-small workspace must be refused before anything is launched */
  if (vda_forward(m, dx, B, T, H, W, dd, ws, ws_bytes / 2, st) == 0) DIE("a half-size workspace was accepted");
  VDA(vda_forward(m, dx, B, T, H, W, dd, ws, ws_bytes, st));
  CU(cudaStreamSynchronize(st));
  float* hd = (float*)malloc(nout * 4);
  CU(cudaMemcpy(hd, dd, nout * 4, cudaMemcpyDeviceToHost));
  f = fopen(argv[10], "wb");
  if (!f || fwrite(hd, 4, nout, f) != nout) DIE("cannot write %s", argv[10]);
  fclose(f);
  printf("ok %d x %d x %d x %d, workspace %lld bytes\n", B, T, H, W, (long long)ws_bytes);
  cudaFree(dx); cudaFree(dd); cudaFree(ws);
  return vda_destroy(m);
}
