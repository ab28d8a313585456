// p2p.cu -- ghost exchange through peer memory (NVLink / NVSwitch) instead of NCCL point-to-point.
//
// What FabArray::FillBoundary's MPI messages are in the reference (a16): here one process per GPU maps every other rank's
// MAILBOX (CUDA IPC) and an exchange is two kernels per rank, no proxy thread and no rendezvous:
//   send kernel : wait for the credit of the slot, store the boundary data straight into the peer's mailbox over NVLink,
//                 __threadfence_system, then raise the peer's data flag (a sequence number) -- the copy and the signal are one launch
//   recv kernel : spin on the own data flag, copy mailbox -> ghost cells, give the credit back to the sender
// Flags are monotonically increasing sequence numbers in the RECEIVER's memory, written by the peer; mailbox slots are double
// buffered (sequence parity), so a sender only ever waits for the consumption of the exchange before the previous one.
// Every rank issues the same exchanges in the same order (the condition NCCL send / recv pairs need as well).
// NCCL remains the transport for all-reduces, for messages larger than a mailbox slot and across nodes.  OPT-IN (IAMRX_P2P=1): see
// p2p_try_exchange for the measurement that keeps NCCL the default.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <unistd.h>
#include "level.h"

#define IX_TRY(call) do { int rc_ = (call); if (rc_ != IAMRX_OK) return rc_; } while (0)

namespace ix {

#if defined(IX_EMUL)
bool p2p_try_exchange(const std::vector<int>&, const std::vector<double*>&, const std::vector<int64_t>&, const std::vector<double*>&,
                      const std::vector<int64_t>&, cudaStream_t, std::vector<char>&, int*) { return false; }
void p2p_finalize() {}
#else

namespace {

constexpr int64_t SLOT_DOUBLES = (int64_t)1 << 20;   // 8 MB per (source rank, parity)
constexpr int MAXMSG = 48;                           // messages per exchange kernel (peers x components x planes)
constexpr int P2P_T = 256, P2P_CHUNKS = 16;

struct Msg {
  const double* src;          // send: local data; recv: own mailbox
  double* dst;                // send: peer mailbox; recv: local ghost region
  int64_t n;
  unsigned long long* wait;   // flag in OWN memory to wait on ...
  unsigned long long wait_for;   // ... until it reaches this value (0: no wait)
  unsigned long long* signal;    // flag in PEER memory raised by the last CTA of the message (nullptr: none)
  unsigned long long signal_val;
  int group;                  // messages of one peer share the wait / signal: counted together
};
struct MsgTable { Msg m[MAXMSG]; int n; int ngroups; int group_size[MAXMSG]; };

__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// grid = (P2P_CHUNKS, nmsg).  done[group] counts finished CTAs of a group; the last one raises the group's signal.
__global__ void __launch_bounds__(P2P_T) p2p_copy_kernel(MsgTable T, unsigned int* done) {
  const Msg& M = T.m[blockIdx.y];
  if (M.wait_for) {
    if (threadIdx.x == 0) while (ld_flag(M.wait) < M.wait_for) { __nanosleep(64); }
    __syncthreads();
  }
  // 16-byte accesses when both ends allow it
  const int64_t n = M.n;
  const bool vec = ((((uintptr_t)M.src) | ((uintptr_t)M.dst)) & 15) == 0;
  const int64_t stride = (int64_t)gridDim.x * P2P_T;
  if (vec) {
    const int64_t n2 = n >> 1;
    const double2* s2 = reinterpret_cast<const double2*>(M.src);
    double2* d2 = reinterpret_cast<double2*>(M.dst);
    for (int64_t i = (int64_t)blockIdx.x * P2P_T + threadIdx.x; i < n2; i += stride) d2[i] = __ldcg(s2 + i);
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) M.dst[n - 1] = __ldcg(M.src + n - 1);
  } else {
    for (int64_t i = (int64_t)blockIdx.x * P2P_T + threadIdx.x; i < n; i += stride) M.dst[i] = __ldcg(M.src + i);
  }
  if (M.signal) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int total = (unsigned int)T.group_size[M.group] * gridDim.x;
      const unsigned int prev = atomicAdd(done + M.group, 1u);
      if (prev + 1 == total) {
        done[M.group] = 0;   // ready for the next launch (stream-ordered)
        __threadfence_system();
        st_flag(M.signal, M.signal_val);
      }
    }
  }
}

struct P2P {
  int state = 0;   // 0 not tried, 1 on, -1 off
  int me = 0, nranks = 1;
  double* mbox = nullptr;                       // [nranks][2][SLOT_DOUBLES]: data FROM rank r, parity
  unsigned long long* flags = nullptr;          // [2][nranks]: data_seq[r] (rank r wrote its message #), cons_seq[r] (rank r consumed mine #)
  std::vector<double*> peer_mbox;
  std::vector<unsigned long long*> peer_flags;
  std::vector<unsigned long long> seq;          // per peer: exchanges done so far
  std::vector<unsigned long long> last_use;     // [peer][parity]: sequence number of my last message in that slot of the peer's mailbox
  unsigned int* done = nullptr;                 // [2][MAXMSG] group counters (send / recv launches)
};
P2P& p2p() { static P2P p; return p; }

void bytes_to_doubles(const void* b, size_t n, double* out) { for (size_t i = 0; i < n; ++i) out[i] = (double)((const unsigned char*)b)[i]; }
void doubles_to_bytes(const double* in, size_t n, void* b) { for (size_t i = 0; i < n; ++i) ((unsigned char*)b)[i] = (unsigned char)in[i]; }

// collective over the NCCL communicator: allocate, exchange IPC handles (as byte-valued doubles through the sum all-reduce),
// map the peers, agree on the outcome
int p2p_setup(cudaStream_t s) {
  P2P& P = p2p();
  Comm& c = comm();
  P.state = -1;
  int want = 1;
  P.me = c.rank; P.nranks = c.nranks;
  const int nr = c.nranks;
  const size_t HB = sizeof(cudaIpcMemHandle_t);
  int ok = want;
  int dev = 0;
  cudaGetDevice(&dev);
  if (ok) {
    if (cudaMalloc(&P.mbox, (size_t)nr * 2 * SLOT_DOUBLES * sizeof(double)) != cudaSuccess) ok = 0;
    if (ok && cudaMalloc(&P.flags, (size_t)2 * nr * sizeof(unsigned long long)) != cudaSuccess) ok = 0;
    if (ok && cudaMalloc(&P.done, 2 * MAXMSG * sizeof(unsigned int)) != cudaSuccess) ok = 0;
    if (ok) { cudaMemset(P.flags, 0, (size_t)2 * nr * sizeof(unsigned long long)); cudaMemset(P.done, 0, 2 * MAXMSG * sizeof(unsigned int)); }
    if (!ok) cudaGetLastError();
  }
  // table: per rank [ok, device, host id hash, 2 handles]
  const size_t per = 3 + 2 * HB;
  std::vector<double> h(per * nr, 0.0);
  cudaIpcMemHandle_t hm{}, hf{};
  if (ok && (cudaIpcGetMemHandle(&hm, P.mbox) != cudaSuccess || cudaIpcGetMemHandle(&hf, P.flags) != cudaSuccess)) { ok = 0; cudaGetLastError(); }
  char host[256] = {0};
  gethostname(host, sizeof(host) - 1);
  unsigned hh = 5381; for (const char* q = host; *q; ++q) hh = hh * 33u + (unsigned char)*q;
  double* mine = h.data() + per * c.rank;
  mine[0] = ok; mine[1] = dev; mine[2] = (double)(hh & 0xffffff);
  bytes_to_doubles(&hm, HB, mine + 3); bytes_to_doubles(&hf, HB, mine + 3 + HB);
  struct G { double* p; ~G() { dev_free(p); } } g{dev_alloc(per * nr)};
  if (!g.p) return IAMRX_ERR_CUDA;
  IX_CUDA(cudaMemcpyAsync(g.p, h.data(), per * nr * sizeof(double), cudaMemcpyHostToDevice, s));
  IX_TRY(comm_allreduce(g.p, (int)(per * nr), 0, s));
  IX_CUDA(cudaMemcpyAsync(h.data(), g.p, per * nr * sizeof(double), cudaMemcpyDeviceToHost, s));
  IX_CUDA(cudaStreamSynchronize(s));
  for (int r = 0; r < nr; ++r) if (h[per * r] != 1.0 || h[per * r + 2] != mine[2]) ok = 0;   // somebody failed, or another node
  P.peer_mbox.assign(nr, nullptr); P.peer_flags.assign(nr, nullptr);
  if (ok) {
    for (int r = 0; r < nr && ok; ++r) {
      if (r == c.rank) { P.peer_mbox[r] = P.mbox; P.peer_flags[r] = P.flags; continue; }
      cudaIpcMemHandle_t a, b;
      doubles_to_bytes(h.data() + per * r + 3, HB, &a); doubles_to_bytes(h.data() + per * r + 3 + HB, HB, &b);
      void* pa = nullptr; void* pb = nullptr;
      if (cudaIpcOpenMemHandle(&pa, a, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          cudaIpcOpenMemHandle(&pb, b, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
      P.peer_mbox[r] = (double*)pa; P.peer_flags[r] = (unsigned long long*)pb;
    }
  }
  // agree: everybody mapped everybody
  double okd = ok;
  IX_CUDA(cudaMemcpyAsync(g.p, &okd, sizeof(double), cudaMemcpyHostToDevice, s));
  IX_TRY(comm_allreduce(g.p, 1, 1, s));
  IX_CUDA(cudaMemcpyAsync(&okd, g.p, sizeof(double), cudaMemcpyDeviceToHost, s));
  IX_CUDA(cudaStreamSynchronize(s));
  P.seq.assign(nr, 0);
  P.last_use.assign((size_t)nr * 2, 0);
  P.state = (okd == 1.0) ? 1 : -1;
  if (getenv("IAMRX_P2P_VERBOSE")) fprintf(stderr, "[iamrx] rank %d: peer-memory exchange %s\n", c.rank, P.state == 1 ? "on" : "off");
  return IAMRX_OK;
}

}  // namespace

// Try to carry the exchange over peer memory.  handled[i] = 1 for the entries that were (the caller sends the rest through NCCL).
// Entries with the same peer are ONE message stream in order (the e-th send to a peer matches its e-th receive from this rank).
// Returns false when the transport is off; *rc carries errors.
bool p2p_try_exchange(const std::vector<int>& peers, const std::vector<double*>& sbuf, const std::vector<int64_t>& scount,
                      const std::vector<double*>& rbuf, const std::vector<int64_t>& rcount, cudaStream_t s, std::vector<char>& handled, int* rc) {
  *rc = IAMRX_OK;
  Comm& c = comm();
  P2P& P = p2p();
  if (c.nranks <= 1 || c.ex || !c.nccl) return false;
  if (P.state == 0) {
    // opt-in: measured on 2 B200s (profiles/r02_notes.md) the peer-memory exchange is no faster than NCCL's fused send / recv kernel for
    // these 1 MB plane exchanges (93.2 vs 93.3 ms per step, slabs) -- what an exchange costs is the pairwise synchronisation of the two
    // ranks, not the transport -- and slower for packed ghost shells (an extra hop through the mailbox: 109.2 vs 101.8 ms, blocks)
    const char* e = getenv("IAMRX_P2P");
    if (!(e && e[0] == '1')) { P.state = -1; return false; }
    *rc = p2p_setup(s);
    if (*rc != IAMRX_OK) return false;
  }
  if (P.state != 1) return false;
  handled.assign(peers.size(), 0);
  // per peer: total sizes decide (identically on both ends) whether the pair uses the mailbox
  std::vector<int64_t> stot(c.nranks, 0), rtot(c.nranks, 0);
  std::vector<int> cnt(c.nranks, 0);
  for (size_t i = 0; i < peers.size(); ++i) { stot[peers[i]] += scount[i]; rtot[peers[i]] += rcount[i]; cnt[peers[i]]++; }
  MsgTable S{}, R{};
  std::vector<int> sgroup(c.nranks, -1), rgroup(c.nranks, -1);
  std::vector<int64_t> soff(c.nranks, 0), roff(c.nranks, 0);
  std::vector<char> use(c.nranks, 0);
  int nmsg_s = 0, nmsg_r = 0;
  for (int p = 0; p < c.nranks; ++p) {
    if (!cnt[p] || p == c.rank) continue;
    if (stot[p] > SLOT_DOUBLES || rtot[p] > SLOT_DOUBLES) continue;
    use[p] = 1;
  }
  for (size_t i = 0; i < peers.size(); ++i) { if (use[peers[i]]) { nmsg_s += scount[i] > 0; nmsg_r += rcount[i] > 0; } }
  if (nmsg_s > MAXMSG || nmsg_r > MAXMSG) return false;   // (all ranks see mirrored counts: the same decision everywhere)
  bool any = false;
  for (int p = 0; p < c.nranks; ++p) if (use[p]) { P.seq[p] += 1; any = true; }
  if (!any) return false;
  for (size_t i = 0; i < peers.size(); ++i) {
    const int p = peers[i];
    if (!use[p]) continue;
    handled[i] = 1;
    const unsigned long long n = P.seq[p];
    const int par = (int)(n & 1);
    if (scount[i] > 0) {
      if (sgroup[p] < 0) { sgroup[p] = S.ngroups++; S.group_size[sgroup[p]] = 0; }
      Msg& m = S.m[S.n++];
      m.src = sbuf[i];
      m.dst = P.peer_mbox[p] + ((int64_t)c.rank * 2 + par) * SLOT_DOUBLES + soff[p];
      m.n = scount[i];
      m.wait = P.flags + c.nranks + p;                 // cons_seq[p] in my memory: p has consumed my message #
      m.wait_for = P.last_use[(size_t)p * 2 + par];    // the slot's previous use (0: never used)
      m.signal = P.peer_flags[p] + c.rank;             // data_seq[me] in p's memory
      m.signal_val = n;
      m.group = sgroup[p];
      S.group_size[m.group]++;
      soff[p] += scount[i];
    }
    if (rcount[i] > 0) {
      if (rgroup[p] < 0) { rgroup[p] = R.ngroups++; R.group_size[rgroup[p]] = 0; }
      Msg& m = R.m[R.n++];
      m.src = P.mbox + ((int64_t)p * 2 + par) * SLOT_DOUBLES + roff[p];
      m.dst = rbuf[i];
      m.n = rcount[i];
      m.wait = P.flags + p;                            // data_seq[p] in my memory
      m.wait_for = n;
      m.signal = P.peer_flags[p] + c.nranks + c.rank;  // cons_seq[me] in p's memory
      m.signal_val = n;
      m.group = rgroup[p];
      R.group_size[m.group]++;
      roff[p] += rcount[i];
    }
  }
  // (a pair with traffic in one direction only: the idle direction neither raises nor waits for a flag -- the counts mirror each other)
  for (int p = 0; p < c.nranks; ++p) if (use[p] && sgroup[p] >= 0) P.last_use[(size_t)p * 2 + (int)(P.seq[p] & 1)] = P.seq[p];
  if (S.n > 0) {
    p2p_copy_kernel<<<dim3(P2P_CHUNKS, S.n), P2P_T, 0, s>>>(S, P.done);
    if (cudaGetLastError() != cudaSuccess) { set_error("p2p send kernel launch failed"); *rc = IAMRX_ERR_CUDA; return true; }
  }
  if (R.n > 0) {
    p2p_copy_kernel<<<dim3(P2P_CHUNKS, R.n), P2P_T, 0, s>>>(R, P.done + MAXMSG);
    if (cudaGetLastError() != cudaSuccess) { set_error("p2p recv kernel launch failed"); *rc = IAMRX_ERR_CUDA; return true; }
  }
  return true;
}

void p2p_finalize() {
  P2P& P = p2p();
  if (P.state == 1) {
    cudaDeviceSynchronize();
    for (int r = 0; r < P.nranks; ++r) {
      if (r == P.me) continue;
      if (P.peer_mbox[r]) cudaIpcCloseMemHandle(P.peer_mbox[r]);
      if (P.peer_flags[r]) cudaIpcCloseMemHandle(P.peer_flags[r]);
    }
  }
  if (P.mbox) cudaFree(P.mbox);
  if (P.flags) cudaFree(P.flags);
  if (P.done) cudaFree(P.done);
  P = P2P{};
}
#endif

}  // namespace ix
