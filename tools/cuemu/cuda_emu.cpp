// cuda_emu.cpp -- see cuda_emu.h.  DEVELOPMENT / TEST INFRASTRUCTURE ONLY.
#include "cuda_emu.h"

namespace emu {
thread_local dim3 t_threadIdx, t_blockIdx;
dim3 g_blockDim, g_gridDim;
Block* g_block = nullptr;

struct ThreadArg { const std::function<void()>* body; dim3 tid, bid; };

static void* thread_main(void* p) {
  ThreadArg* a = (ThreadArg*)p;
  t_threadIdx = a->tid;
  t_blockIdx = a->bid;
  (*a->body)();
  return nullptr;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const unsigned nthr = block.x * block.y * block.z;
  if (nthr == 0 || grid.x * grid.y * grid.z == 0) return;
  if (block.y != 1 || block.z != 1) { fprintf(stderr, "emu: 1-D blocks only\n"); abort(); }
  g_blockDim = block; g_gridDim = grid;
  Block blk;
  g_block = &blk;
  // heap allocation so AddressSanitizer sees out-of-bounds shared-memory accesses
  blk.dyn = (unsigned char*)aligned_alloc(128, ((smem + 127) / 128 + 1) * 128);
  std::vector<pthread_t> th(nthr);
  std::vector<ThreadArg> args(nthr);
  pthread_attr_t attr; pthread_attr_init(&attr); pthread_attr_setstacksize(&attr, 1 << 20);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        pthread_barrier_init(&blk.bar, nullptr, nthr);
        const unsigned nwarps = (nthr + 31) / 32;
        for (unsigned w = 0; w < nwarps; ++w)
          pthread_barrier_init(&blk.warp_bar[w], nullptr, std::min(32u, nthr - w * 32));
        memset(blk.dyn, 0xCB, smem);   // poison: uninitialised shared reads show up as garbage
        for (unsigned t = 0; t < nthr; ++t) {
          args[t].body = &body; args[t].tid = dim3(t, 0, 0); args[t].bid = dim3(bx, by, bz);
          if (pthread_create(&th[t], &attr, thread_main, &args[t]) != 0) { perror("pthread_create"); abort(); }
        }
        for (unsigned t = 0; t < nthr; ++t) pthread_join(th[t], nullptr);
        pthread_barrier_destroy(&blk.bar);
        for (unsigned w = 0; w < nwarps; ++w) pthread_barrier_destroy(&blk.warp_bar[w]);
      }
  free(blk.dyn);
  g_block = nullptr;
}
}  // namespace emu
