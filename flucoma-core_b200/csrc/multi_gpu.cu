// Product-level multi-GPU entry (north_star: "partition across the box's GPUs with one NCCL allgather only for
// the final activations").  One process, one host thread per device: the batch is cut into contiguous shards
// (fb200_shard_range), every shard runs the complete BufNMF pipeline on its own plan, nothing is exchanged on the data
// path.  Optionally the final activations of ALL buffers are gathered onto every device with ONE ncclAllGather.
// NCCL is loaded lazily with dlopen (libnccl.so.2), so hosts without it can still use everything else.
#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <atomic>
#include <mutex>
#include <thread>

using namespace fb200;
struct fb200_plan : public fb200::Plan {};

namespace {

struct Nccl {
  void* handle = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string& err)
  {
    if (handle) return true;
    handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!handle) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
    auto sym = [&](const char* n) { return dlsym(handle, n); };
    CommInitAll = reinterpret_cast<decltype(CommInitAll)>(sym("ncclCommInitAll"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
    GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
    GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
    AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
    if (!CommInitAll || !CommDestroy || !GroupStart || !GroupEnd || !AllGather || !GetErrorString) {
      err = "libnccl.so.2 lacks the expected symbols";
      return false;
    }
    return true;
  }
};

struct CommCache { // communicators are expensive to build: one set per distinct device list, kept for the process
  std::mutex mu;
  Nccl nccl;
  std::map<std::vector<int>, std::vector<ncclComm_t>> comms;
};
CommCache& cache()
{
  static CommCache c;
  return c;
}

// progress of a sharded job: the callback sees the minimum over the shards, each iteration exactly once, from one thread
// at a time; a cancel is sticky for every shard
struct SharedProgress {
  fb200_progress_fn fn;
  void* user;
  std::mutex mu;
  std::vector<int64_t> at;
  int64_t reported = 0;
  std::atomic<bool> cancelled{false};
};
struct ShardProgress {
  SharedProgress* sp;
  int shard;
};
int shard_progress(void* u, int64_t it)
{
  auto* s = static_cast<ShardProgress*>(u);
  SharedProgress& sp = *s->sp;
  if (sp.cancelled.load()) return 0;
  std::lock_guard<std::mutex> lock(sp.mu);
  sp.at[(size_t) s->shard] = it;
  int64_t lo = it;
  for (int64_t v : sp.at) lo = std::min(lo, v);
  while (sp.reported < lo && !sp.cancelled.load())
    if (!sp.fn(sp.user, ++sp.reported)) sp.cancelled.store(true);
  return sp.cancelled.load() ? 0 : 1;
}

} // namespace

extern "C" int32_t fb200_bufnmf_sharded(const fb200_sharded_args* a)
{
  if (!a || a->struct_size != sizeof(fb200_sharded_args) || a->n_devices <= 0 || !a->plans || !a->job ||
      a->job->struct_size != sizeof(fb200_bufnmf_args))
    return FB200_ERR_INVALID;
  const fb200_bufnmf_args& J = *a->job;
  const int world = a->n_devices;
  for (int r = 0; r < world; r++)
    if (!a->plans[r]) return FB200_ERR_INVALID;
  fb200_plan* p0 = a->plans[0];
  if (J.mem != FB200_HOST) { p0->err = "fb200_bufnmf_sharded: the job's arrays must be host memory"; return FB200_ERR_INVALID; }
  if (J.batch <= 0 || J.n_samples <= 0 || J.rank <= 0) { p0->err = "fb200_bufnmf_sharded: bad job"; return FB200_ERR_INVALID; }
  for (int r = 0; r < world; r++) {
    if (a->plans[r]->win != p0->win || a->plans[r]->hop != p0->hop || a->plans[r]->fft != p0->fft) {
      p0->err = "fb200_bufnmf_sharded: the plans differ in their FFT settings";
      return FB200_ERR_INVALID;
    }
    for (int q = 0; q < r; q++)
      if (a->plans[q]->cfg.device == a->plans[r]->cfg.device) { p0->err = "fb200_bufnmf_sharded: two plans on one device"; return FB200_ERR_INVALID; }
  }
  const int64_t n = J.n_samples, K = J.rank, B = p0->bins;
  const int64_t F = fb200_num_frames(n, p0->win, p0->hop);
  const bool gather = a->gathered_acts != nullptr && J.acts_out != nullptr && J.acts_mode != 2;

  std::vector<fb200_bufnmf_args> jobs((size_t) world, J);
  std::vector<int64_t> begin((size_t) world), count((size_t) world);
  SharedProgress sp{J.progress, J.progress_user};
  sp.at.assign((size_t) world, 0);
  std::vector<ShardProgress> spu((size_t) world);
  for (int r = 0; r < world; r++) {
    fb200_shard_range(J.batch, world, r, &begin[(size_t) r], &count[(size_t) r]);
    fb200_bufnmf_args& s = jobs[(size_t) r];
    const int64_t b0 = begin[(size_t) r];
    s.batch = count[(size_t) r];
    s.audio = J.audio + b0 * n;
    s.seeds = J.seeds ? J.seeds + b0 : nullptr;
    s.bases_in = J.bases_in ? J.bases_in + b0 * K * B : nullptr;
    s.acts_in = J.acts_in ? J.acts_in + b0 * F * K : nullptr;
    s.bases_out = J.bases_out ? J.bases_out + b0 * K * B : nullptr;
    s.acts_out = J.acts_out ? J.acts_out + b0 * F * K : nullptr;
    s.resynth_out = J.resynth_out ? J.resynth_out + b0 * K * n : nullptr;
    if (J.progress) {
      spu[(size_t) r] = ShardProgress{&sp, r};
      s.progress = &shard_progress;
      s.progress_user = &spu[(size_t) r];
      if (count[(size_t) r] == 0) sp.at[(size_t) r] = INT64_MAX; // an empty shard never holds the others back
    }
  }
  std::vector<int32_t> status((size_t) world, FB200_OK);
  {
    std::vector<std::thread> th;
    for (int r = 0; r < world; r++)
      th.emplace_back([&, r]() {
        if (count[(size_t) r] > 0) status[(size_t) r] = fb200_bufnmf(a->plans[r], &jobs[(size_t) r]);
      });
    for (auto& t : th) t.join();
  }
  int32_t worst = FB200_OK;
  for (int r = 0; r < world; r++) {
    const int32_t s = status[(size_t) r];
    if (s < 0) {
      if (r != 0) p0->err = "shard " + std::to_string(r) + ": " + a->plans[r]->err;
      return s;
    }
    if (s == FB200_CANCELLED) worst = FB200_CANCELLED;
    else if (s > worst && worst != FB200_CANCELLED) worst = s;
  }
  if (worst == FB200_CANCELLED || !gather) return worst;

  // ---- one all-gather of the final (scaled) activations: every device ends up with all of them --------------------------
  // shards are padded to the largest one so that the counts are equal (ncclAllGather): shard r occupies rows
  // [r * per, r * per + count_r) of gathered_acts[d], per = ceil(batch / world)
  CommCache& cc = cache();
  std::lock_guard<std::mutex> lock(cc.mu);
  if (!cc.nccl.load(p0->err)) return FB200_ERR_UNSUPPORTED;
  std::vector<int> devs((size_t) world);
  for (int r = 0; r < world; r++) devs[(size_t) r] = a->plans[r]->cfg.device;
  auto it = cc.comms.find(devs);
  if (it == cc.comms.end()) {
    std::vector<ncclComm_t> comms((size_t) world);
    ncclResult_t rc = cc.nccl.CommInitAll(comms.data(), world, devs.data());
    if (rc != ncclSuccess) { p0->err = std::string("ncclCommInitAll: ") + cc.nccl.GetErrorString(rc); return FB200_ERR_CUDA; }
    it = cc.comms.emplace(devs, std::move(comms)).first;
  }
  const int64_t per = (J.batch + world - 1) / world;
  const size_t elems = (size_t) (per * F * K);
  for (int r = 0; r < world; r++) { // the send buffer: this shard's activations, still on the device from the export
    Plan* p = a->plans[r];
    FB_CUDA(p, cudaSetDevice(p->cfg.device));
    // grows only for a short shard (rows beyond count_r are padding); the shard's activations are kept
    FB_CUDA(p, p->out_b.ensure_keep(sizeof(float) * elems, sizeof(float) * (size_t) (count[(size_t) r] * F * K)));
  }
  ncclResult_t rc = cc.nccl.GroupStart();
  for (int r = 0; r < world && rc == ncclSuccess; r++) {
    Plan* p = a->plans[r];
    cudaSetDevice(p->cfg.device);
    rc = cc.nccl.AllGather(p->out_b.p, a->gathered_acts[r], elems, ncclFloat, it->second[(size_t) r], p->stream);
  }
  ncclResult_t rc2 = cc.nccl.GroupEnd();
  if (rc != ncclSuccess || rc2 != ncclSuccess) {
    p0->err = std::string("ncclAllGather: ") + cc.nccl.GetErrorString(rc != ncclSuccess ? rc : rc2);
    return FB200_ERR_CUDA;
  }
  for (int r = 0; r < world; r++) {
    Plan* p = a->plans[r];
    FB_CUDA(p, cudaSetDevice(p->cfg.device));
    FB_CUDA(p, cudaStreamSynchronize(p->stream));
  }
  return worst;
}
