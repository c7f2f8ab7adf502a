// comm.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink / NVSwitch.
// The path shards by points: every rank owns a contiguous point range and its
// observations; the only exchanges are FP64 sum all-reduces of the replicated camera-side
// quantities (U, ga once per linearisation; the S tile pool and ea once per damping try; a few
// scalars per try), issued on the same stream as the kernels so no host synchronisation is added.
#include "psba_internal.h"
#include <nccl.h>
#include <string.h>

#define NCCL_CHECK(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) { \
    fprintf(stderr, "psba_b200: NCCL error %d (%s) at %s(%d)\n", (int)r_, ncclGetErrorString(r_), __FILE__, __LINE__); \
    exit(EXIT_FAILURE); } } while (0)

static ncclComm_t g_comm;
static bool g_active = false;
static int g_rank = 0, g_size = 1;

bool psba_comm_active() { return g_active; }
int psba_comm_rank() { return g_rank; }
int psba_comm_size() { return g_size; }

extern "C" void psba_comm_unique_id(char *out128)
{
    ncclUniqueId id;
    NCCL_CHECK(ncclGetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
}

extern "C" void psba_comm_init(int rank, int nranks, const char *unique_id128)
{
    if (g_active) return;
    if (nranks <= 1) { g_rank = 0; g_size = 1; return; }
    ncclUniqueId id;
    memcpy(&id, unique_id128, 128);
    NCCL_CHECK(ncclCommInitRank(&g_comm, nranks, id, rank));
    g_active = true; g_rank = rank; g_size = nranks;
}

extern "C" void psba_comm_finalize(void)
{
    if (g_active) { ncclCommDestroy(g_comm); g_active = false; g_rank = 0; g_size = 1; }
}

void psba_allreduce_sum(psba_ctx *c, double *buf, size_t count)
{
    if (!g_active) return;
    PROF(c, KID_ALLREDUCE) NCCL_CHECK(ncclAllReduce(buf, buf, count, ncclDouble, ncclSum, g_comm, c->stream));
}

void psba_allreduce_max(psba_ctx *c, double *buf, size_t count)
{
    if (!g_active) return;
    NCCL_CHECK(ncclAllReduce(buf, buf, count, ncclDouble, ncclMax, g_comm, c->stream));
}
