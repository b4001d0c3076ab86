// Inter-GPU communication of the Krylov solve: the all-reduce behind every
// Gram–Schmidt dot / norm and the ghost-dof halo exchange before every SpMV.
//
// Replaces what the reference does through Epetra over MPI (SURVEY.md §2
// "Parallelism and communication": Epetra_Import of ghost values, MPI_Allreduce in
// AztecOO's dots).  One NCCL communicator per context, one rank per GPU; NCCL is
// loaded at run time (dlopen) so a single-GPU process needs no NCCL at all.
#include <dlfcn.h>
#include <string.h>

#include "context.h"

namespace glsns
{
  namespace
  {
    struct Uid
    {
      char internal[128];
    };
    struct NcclApi
    {
      void *handle = nullptr;
      int (*GetUniqueId)(void *)                                                   = nullptr;
      int (*CommInitRank)(void **, int, /*ncclUniqueId by value*/ Uid, int) = nullptr;
      int (*CommDestroy)(void *)                                                   = nullptr;
      int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
      int (*Send)(const void *, size_t, int, int, void *, cudaStream_t)            = nullptr;
      int (*Recv)(void *, size_t, int, int, void *, cudaStream_t)                  = nullptr;
      int (*GroupStart)()                                                           = nullptr;
      int (*GroupEnd)()                                                             = nullptr;
      const char *(*GetErrorString)(int)                                            = nullptr;
    };
    constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2;

    NcclApi &
    api()
    {
      static NcclApi a;
      return a;
    }

    bool
    load_nccl(std::string &why)
    {
      NcclApi &a = api();
      if (a.handle)
        return true;
      const char *names[] = {"libnccl.so.2", "libnccl.so"};
      for (const char *nm : names)
        if ((a.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL)))
          break;
      if (!a.handle)
        {
          why = std::string("cannot load NCCL: ") + dlerror();
          return false;
        }
#define SYM(field, name)                                        \
  *(void **)(&a.field) = dlsym(a.handle, name);                 \
  if (!a.field)                                                 \
    {                                                           \
      why = std::string("NCCL symbol missing: ") + name;        \
      return false;                                             \
    }
      SYM(GetUniqueId, "ncclGetUniqueId")
      SYM(CommInitRank, "ncclCommInitRank")
      SYM(CommDestroy, "ncclCommDestroy")
      SYM(AllReduce, "ncclAllReduce")
      SYM(Send, "ncclSend")
      SYM(Recv, "ncclRecv")
      SYM(GroupStart, "ncclGroupStart")
      SYM(GroupEnd, "ncclGroupEnd")
      SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
      return true;
    }

    glsns_status
    nccl_check(glsns_context *ctx, int rc, const char *what)
    {
      if (rc == 0)
        return GLSNS_OK;
      return fail(ctx, GLSNS_ERR_COMM, std::string(what) + ": " + api().GetErrorString(rc));
    }

    __global__ void
    pack_kernel(const int64_t n, const int32_t *__restrict__ idx, const double *__restrict__ x,
                double *__restrict__ buf)
    {
      const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (t < n)
        buf[t] = x[idx[t]];
    }
  } // namespace

  glsns_status
  comm_unique_id(uint8_t out[128])
  {
    std::string why;
    if (!load_nccl(why))
      return GLSNS_ERR_COMM;
    Uid id;
    if (api().GetUniqueId(&id) != 0)
      return GLSNS_ERR_COMM;
    memcpy(out, id.internal, 128);
    return GLSNS_OK;
  }

  glsns_status
  comm_init(glsns_context *ctx, int32_t n_ranks, int32_t rank, const uint8_t unique_id[128])
  {
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks)
      return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "bad rank / n_ranks");
    ctx->n_ranks = n_ranks;
    ctx->rank    = rank;
    if (n_ranks == 1)
      return GLSNS_OK;
    std::string why;
    if (!load_nccl(why))
      return fail(ctx, GLSNS_ERR_COMM, why);
    Uid id;
    memcpy(id.internal, unique_id, 128);
    return nccl_check(ctx, api().CommInitRank(&ctx->nccl_comm, n_ranks, id, rank),
                      "ncclCommInitRank");
  }

  void
  comm_destroy(glsns_context *ctx)
  {
    if (ctx->nccl_comm)
      api().CommDestroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }

  glsns_status
  allreduce_sum(glsns_context *ctx, double *dev, int n)
  {
    if (ctx->n_ranks == 1)
      return GLSNS_OK;
    return nccl_check(ctx,
                      api().AllReduce(dev, dev, (size_t)n, NCCL_FLOAT64, NCCL_SUM, ctx->nccl_comm,
                                      ctx->stream),
                      "ncclAllReduce");
  }

  glsns_status
  allreduce_max(glsns_context *ctx, double *dev, int n)
  {
    if (ctx->n_ranks == 1)
      return GLSNS_OK;
    return nccl_check(ctx,
                      api().AllReduce(dev, dev, (size_t)n, NCCL_FLOAT64, NCCL_MAX, ctx->nccl_comm,
                                      ctx->stream),
                      "ncclAllReduce");
  }

  // ghosted[n_owned + recv range of neighbour i] <- neighbour i's owned values
  glsns_status
  halo_exchange(glsns_context *ctx, double *ghosted)
  {
    if (ctx->n_ranks == 1 || ctx->n_neighbors == 0)
      return GLSNS_OK;
    const int64_t n_send = ctx->send_ptr[ctx->n_neighbors];
    if (n_send)
      {
        pack_kernel<<<(unsigned)((n_send + 255) / 256), 256, 0, ctx->stream>>>(
          n_send, ctx->send_idx.p, ghosted, ctx->send_buf.p);
        ctx->kernel_launches++;
      }
    GLSNS_TRY(nccl_check(ctx, api().GroupStart(), "ncclGroupStart"));
    for (int i = 0; i < ctx->n_neighbors; ++i)
      {
        const int64_t ns = ctx->send_ptr[i + 1] - ctx->send_ptr[i];
        const int64_t nr = ctx->recv_ptr[i + 1] - ctx->recv_ptr[i];
        if (ns)
          GLSNS_TRY(nccl_check(ctx,
                               api().Send(ctx->send_buf.p + ctx->send_ptr[i], (size_t)ns,
                                          NCCL_FLOAT64, ctx->neighbor_rank[i], ctx->nccl_comm,
                                          ctx->stream),
                               "ncclSend"));
        if (nr)
          GLSNS_TRY(nccl_check(ctx,
                               api().Recv(ghosted + ctx->n_owned + ctx->recv_ptr[i], (size_t)nr,
                                          NCCL_FLOAT64, ctx->neighbor_rank[i], ctx->nccl_comm,
                                          ctx->stream),
                               "ncclRecv"));
      }
    return nccl_check(ctx, api().GroupEnd(), "ncclGroupEnd");
  }
} // namespace glsns
