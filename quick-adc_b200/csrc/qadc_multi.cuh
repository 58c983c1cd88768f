// qadc_multi.cuh — one database sharded over the GPUs of one box, driven by ONE host process/thread
// (included at the end of qadc_capi.cu, same translation unit).
//
// A qadc_multi owns one qadc_ctx per device.  Flat databases are cut into contiguous runs of whole
// 256-vector superblocks (global positions kept through qadc_set_position_base, so the canonical order
// is preserved); inverted lists are dealt whole to the devices, longest first; every shard holds a
// replica of all keep-prefixes, so all shards derive bit-identical qmin / qmax / int8 tables without a
// collective.  Exchange steps per query batch: [inverted lists] one all-gather of the per-shard coarse
// candidates (the K cells are split over the devices, K/G per GPU), then one all-gather of the per-shard
// top-r (key, id) lists, merged on device 0 under the same total order -> the result equals the
// single-GPU one bit for bit.  The all-gathers are NCCL (ncclCommInitAll, one communicator per device,
// grouped calls on the contexts' streams; libnccl.so.2 is opened at run time).  When a device ordinal is
// listed more than once (virtual shards on one GPU: how a single-GPU test box exercises this layer) NCCL
// cannot be used — it rejects duplicate devices — and the gather is G same-device copies instead.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <numeric>

struct qadc_multi {
    int G = 0;
    std::vector<int> devices;
    std::vector<qadc_ctx*> ctx;
    std::string err;
    bool use_nccl = false;
    std::vector<ncclComm_t> comms;
    // NCCL entry points (dlopen)
    void* nccl_lib = nullptr;
    ncclResult_t (*pCommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*pCommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*pAllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*pGroupStart)() = nullptr;
    ncclResult_t (*pGroupEnd)() = nullptr;
    const char* (*pGetErrorString)(ncclResult_t) = nullptr;
    // database description
    int dim = 0, m = 0, K = 0, parts = 0;
    bool loaded = false;
    // per-device scratch
    struct Shard {
        DevBuf q, ids, dists, counts, keys, gkeys, gids, assign, ckeys, gckeys, oids, odists, ocounts, local, glocal;
    };
    std::vector<Shard> sh;
    cudaEvent_t ev[6] = {};   // on device 0's stream: start, coarse done, search done, gather done, merge done, d2h done
};

namespace {

int mfail(qadc_multi* m, int code, const std::string& msg) {
    if (m) m->err = msg;
    else g_create_error = msg;
    return code;
}
#define MCK(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess) return mfail(mm, QADC_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)
#define MCTX(g, call)                                                                                      \
    do {                                                                                                   \
        int rc__ = (call);                                                                                 \
        if (rc__) return mfail(mm, rc__, std::string("device ") + std::to_string(mm->devices[g]) + ": " + qadc_last_error(mm->ctx[g])); \
    } while (0)
#define MNCCL(call)                                                                                        \
    do {                                                                                                   \
        ncclResult_t r__ = (call);                                                                         \
        if (r__ != ncclSuccess) return mfail(mm, QADC_ECUDA, std::string(#call) + ": " + mm->pGetErrorString(r__)); \
    } while (0)

int mensure(qadc_multi* mm, int g, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return QADC_OK;
    MCK(cudaSetDevice(mm->devices[g]));
    if (b.p) MCK(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    MCK(cudaMalloc(&b.p, bytes + bytes / 8 + 256));
    b.cap = bytes + bytes / 8 + 256;
    return QADC_OK;
}

// all-gathers of `bytes` per shard: send[g] -> recv[g] laid out [G][bytes] on every device; all jobs of one call travel
// in ONE NCCL group (a single fused launch per device)
struct GatherJob {
    std::vector<const void*> send;
    std::vector<void*> recv;
    size_t bytes;
};
int multi_all_gather(qadc_multi* mm, const std::vector<GatherJob>& jobs) {
    if (mm->use_nccl) {
        MNCCL(mm->pGroupStart());
        for (const GatherJob& j : jobs)
            for (int g = 0; g < mm->G; ++g) {
                ncclResult_t r = mm->pAllGather(j.send[g], j.recv[g], j.bytes, ncclUint8, mm->comms[g], mm->ctx[g]->stream);
                if (r != ncclSuccess) { mm->pGroupEnd(); return mfail(mm, QADC_ECUDA, std::string("ncclAllGather: ") + mm->pGetErrorString(r)); }
            }
        MNCCL(mm->pGroupEnd());
        return QADC_OK;
    }
    // virtual shards on one device: every source stream must have produced its buffer before anyone copies it
    for (int g = 0; g < mm->G; ++g) MCK(cudaStreamSynchronize(mm->ctx[g]->stream));
    for (const GatherJob& j : jobs)
        for (int g = 0; g < mm->G; ++g)
            for (int s = 0; s < mm->G; ++s)
                MCK(cudaMemcpyAsync(static_cast<uint8_t*>(j.recv[g]) + static_cast<size_t>(s) * j.bytes, j.send[s], j.bytes,
                                    cudaMemcpyDeviceToDevice, mm->ctx[g]->stream));
    return QADC_OK;
}
int multi_all_gather(qadc_multi* mm, const std::vector<const void*>& send, const std::vector<void*>& recv, size_t bytes) {
    return multi_all_gather(mm, std::vector<GatherJob>{GatherJob{send, recv, bytes}});
}

}  // namespace

extern "C" {

int qadc_multi_create(const int* devices, int n, qadc_multi** out) {
    qadc_multi* mm = nullptr;
    if (!devices || n <= 0 || !out) return mfail(nullptr, QADC_EINVAL, "bad device list");
    mm = new qadc_multi;
    mm->G = n;
    mm->devices.assign(devices, devices + n);
    mm->ctx.assign(n, nullptr);
    mm->sh.resize(n);
    auto bail = [&](int rc) {
        g_create_error = mm->err;
        qadc_multi_destroy(mm);
        return rc;
    };
    for (int g = 0; g < n; ++g) {
        int rc = qadc_create(devices[g], nullptr, &mm->ctx[g]);
        if (rc) { mm->err = g_create_error; return bail(rc); }
    }
    std::vector<int> sorted(mm->devices);
    std::sort(sorted.begin(), sorted.end());
    const bool distinct = std::adjacent_find(sorted.begin(), sorted.end()) == sorted.end();
    if (n > 1 && distinct) {
        mm->nccl_lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!mm->nccl_lib) { mm->err = std::string("cannot open libnccl.so.2: ") + dlerror(); return bail(QADC_ECUDA); }
        mm->pCommInitAll = reinterpret_cast<decltype(mm->pCommInitAll)>(dlsym(mm->nccl_lib, "ncclCommInitAll"));
        mm->pCommDestroy = reinterpret_cast<decltype(mm->pCommDestroy)>(dlsym(mm->nccl_lib, "ncclCommDestroy"));
        mm->pAllGather = reinterpret_cast<decltype(mm->pAllGather)>(dlsym(mm->nccl_lib, "ncclAllGather"));
        mm->pGroupStart = reinterpret_cast<decltype(mm->pGroupStart)>(dlsym(mm->nccl_lib, "ncclGroupStart"));
        mm->pGroupEnd = reinterpret_cast<decltype(mm->pGroupEnd)>(dlsym(mm->nccl_lib, "ncclGroupEnd"));
        mm->pGetErrorString = reinterpret_cast<decltype(mm->pGetErrorString)>(dlsym(mm->nccl_lib, "ncclGetErrorString"));
        if (!mm->pCommInitAll || !mm->pCommDestroy || !mm->pAllGather || !mm->pGroupStart || !mm->pGroupEnd || !mm->pGetErrorString) {
            mm->err = "libnccl.so.2 lacks a required entry point";
            return bail(QADC_ECUDA);
        }
        mm->comms.assign(n, nullptr);
        ncclResult_t r = mm->pCommInitAll(mm->comms.data(), n, mm->devices.data());
        if (r != ncclSuccess) { mm->err = std::string("ncclCommInitAll: ") + mm->pGetErrorString(r); mm->comms.clear(); return bail(QADC_ECUDA); }
        mm->use_nccl = true;
    }
    cudaSetDevice(mm->devices[0]);
    for (auto& e : mm->ev)
        if (cudaEventCreate(&e) != cudaSuccess) { mm->err = "cudaEventCreate failed"; return bail(QADC_ECUDA); }
    *out = mm;
    return QADC_OK;
}

void qadc_multi_destroy(qadc_multi* mm) {
    if (!mm) return;
    for (int g = 0; g < mm->G; ++g) {
        if (!mm->ctx[g]) continue;
        cudaSetDevice(mm->devices[g]);
        cudaStreamSynchronize(mm->ctx[g]->stream);
        for (DevBuf* b : {&mm->sh[g].q, &mm->sh[g].ids, &mm->sh[g].dists, &mm->sh[g].counts, &mm->sh[g].keys, &mm->sh[g].gkeys,
                          &mm->sh[g].gids, &mm->sh[g].assign, &mm->sh[g].ckeys, &mm->sh[g].gckeys, &mm->sh[g].oids,
                          &mm->sh[g].odists, &mm->sh[g].ocounts, &mm->sh[g].local, &mm->sh[g].glocal})
            cudaFree(b->p);
    }
    if (mm->use_nccl)
        for (auto c : mm->comms)
            if (c) mm->pCommDestroy(c);
    if (mm->G) cudaSetDevice(mm->devices[0]);
    for (auto& e : mm->ev)
        if (e) cudaEventDestroy(e);
    for (auto c : mm->ctx) qadc_destroy(c);
    // the NCCL library stays loaded (dlclose of a CUDA library at exit is not worth the risk)
    delete mm;
}

const char* qadc_multi_last_error(const qadc_multi* mm) { return mm ? mm->err.c_str() : g_create_error.c_str(); }
int qadc_multi_device_count(const qadc_multi* mm) { return mm ? mm->G : 0; }
int qadc_multi_uses_nccl(const qadc_multi* mm) { return mm && mm->use_nccl ? 1 : 0; }
qadc_ctx* qadc_multi_context(qadc_multi* mm, int g) { return (mm && g >= 0 && g < mm->G) ? mm->ctx[g] : nullptr; }

int qadc_multi_set_pq(qadc_multi* mm, int dim, int m, int bits, const float* codebooks, const float* rotation) {
    if (!mm) return QADC_EINVAL;
    for (int g = 0; g < mm->G; ++g) MCTX(g, qadc_set_pq(mm->ctx[g], dim, m, bits, codebooks, rotation));
    mm->dim = dim; mm->m = m;
    return QADC_OK;
}

int qadc_multi_set_coarse(qadc_multi* mm, int K, const float* centroids) {
    if (!mm) return QADC_EINVAL;
    for (int g = 0; g < mm->G; ++g) MCTX(g, qadc_set_coarse(mm->ctx[g], K, centroids));
    mm->K = K > 0 ? K : 0;
    return QADC_OK;
}

int qadc_multi_load(qadc_multi* mm, int partition_count, const uint32_t* sizes, const uint8_t* const* part_codes,
                    const uint32_t* const* part_labels, float keep) {
    if (!mm || !sizes || !part_codes || partition_count <= 0) return mfail(mm, QADC_EINVAL, "bad database description");
    if (mm->m == 0) return mfail(mm, QADC_ESTATE, "qadc_multi_set_pq must be called first");
    const int G = mm->G, P = partition_count;
    const size_t CS = mm->m / 2;
    auto start_size = [&](uint32_t size) -> uint32_t {   // db_query_4.cpp:125-126, float32 arithmetic
        if (size == 0) return 0u;
        return std::min(size, std::max(1u, static_cast<unsigned>(static_cast<float>(size) * keep)));
    };
    if (mm->K == 0) {
        if (P != 1) return mfail(mm, QADC_EINVAL, "a flat database has exactly one partition");
        const uint64_t n = sizes[0];
        const uint64_t n_sb = (n + kSbVec - 1) / kSbVec;
        const uint32_t npre = start_size(sizes[0]);
        for (int g = 0; g < G; ++g) {
            const uint64_t base = n_sb / G, extra = n_sb % G;
            const uint64_t sb_lo = g * base + std::min<uint64_t>(g, extra), sb_hi = sb_lo + base + (static_cast<uint64_t>(g) < extra ? 1 : 0);
            const uint64_t lo = std::min(sb_lo * kSbVec, n), hi = std::min(sb_hi * kSbVec, n);
            const uint32_t n_local = static_cast<uint32_t>(hi - lo);
            MCTX(g, qadc_begin_database(mm->ctx[g], 1, &n_local, part_labels && part_labels[0] ? 1 : 0));
            MCTX(g, qadc_upload_codes(mm->ctx[g], 0, 0, n_local, part_codes[0] + lo * CS,
                                      part_labels && part_labels[0] ? part_labels[0] + lo : nullptr, 0));
            if (G > 1) {
                MCTX(g, qadc_set_position_base(mm->ctx[g], 0, static_cast<uint32_t>(lo)));
                if (npre) MCTX(g, qadc_set_prefix(mm->ctx[g], 0, part_codes[0], npre, 0));
            }
            MCTX(g, qadc_finalize(mm->ctx[g], keep));
        }
    } else {
        if (P != mm->K) return mfail(mm, QADC_EINVAL, "partition_count != coarse centroid count");
        if (!part_labels) return mfail(mm, QADC_EINVAL, "inverted lists need labels");
        // whole lists, longest first, each to the least loaded device
        std::vector<int> order(P), owner(P, 0);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return sizes[a] > sizes[b]; });
        std::vector<uint64_t> load(G, 0);
        for (int p : order) {
            const int g = static_cast<int>(std::min_element(load.begin(), load.end()) - load.begin());
            owner[p] = g;
            load[g] += sizes[p];
        }
        // no prefix replicas: the table pipeline is "owner computes" (qadc_tables_local_device / qadc_search_bounded_device)
        std::vector<uint32_t> lsizes(P);
        std::vector<const uint8_t*> lcodes(P);
        std::vector<const uint32_t*> llabels(P);
        for (int g = 0; g < G; ++g) {
            for (int p = 0; p < P; ++p) {
                const bool mine = owner[p] == g;
                lsizes[p] = mine ? sizes[p] : 0u;
                lcodes[p] = mine ? part_codes[p] : nullptr;
                llabels[p] = mine ? part_labels[p] : nullptr;
            }
            MCTX(g, qadc_begin_database(mm->ctx[g], P, lsizes.data(), 1));
            if (G > 1) {   // empty lists too have exactly one owner (their tables take part in qmin)
                std::vector<uint8_t> mask(P);
                for (int p = 0; p < P; ++p) mask[p] = owner[p] == g ? 1 : 0;
                MCTX(g, qadc_set_owned_partitions(mm->ctx[g], mask.data()));
            }
            MCTX(g, qadc_upload_partitions(mm->ctx[g], lcodes.data(), llabels.data()));
            MCTX(g, qadc_finalize(mm->ctx[g], keep));
        }
    }
    mm->parts = P;
    mm->loaded = true;
    return QADC_OK;
}

int qadc_multi_search(qadc_multi* mm, const float* queries, int nq, int ma, int r, uint32_t* out_ids, int8_t* out_dists,
                      int32_t* out_counts, qadc_metrics* metrics) {
    if (!mm) return QADC_EINVAL;
    if (!mm->loaded) return mfail(mm, QADC_ESTATE, "no database loaded");
    if (!queries || !out_ids || !out_dists || nq <= 0) return mfail(mm, QADC_EINVAL, "bad arguments");
    const int G = mm->G;
    const size_t nr = static_cast<size_t>(nq) * r, qbytes = static_cast<size_t>(nq) * mm->dim * 4;
    const bool ivf = mm->K > 0;
    for (int g = 0; g < G; ++g) {
        auto& s = mm->sh[g];
        int rc = 0;
        rc = rc ? rc : mensure(mm, g, s.q, qbytes);
        rc = rc ? rc : mensure(mm, g, s.ids, nr * 4);
        rc = rc ? rc : mensure(mm, g, s.dists, nr);
        rc = rc ? rc : mensure(mm, g, s.counts, static_cast<size_t>(nq) * 4);
        rc = rc ? rc : mensure(mm, g, s.keys, nr * 8);
        if (G > 1) {
            rc = rc ? rc : mensure(mm, g, s.gkeys, nr * 8 * G);
            if (ivf) {
                rc = rc ? rc : mensure(mm, g, s.gids, nr * 4 * G);
                rc = rc ? rc : mensure(mm, g, s.assign, static_cast<size_t>(nq) * ma * 4);
                rc = rc ? rc : mensure(mm, g, s.ckeys, static_cast<size_t>(nq) * ma * 8);
                rc = rc ? rc : mensure(mm, g, s.gckeys, static_cast<size_t>(nq) * ma * 8 * G);
                rc = rc ? rc : mensure(mm, g, s.local, static_cast<size_t>(std::min(nq, 32768)) * (r + 1) * 4);
                rc = rc ? rc : mensure(mm, g, s.glocal, static_cast<size_t>(std::min(nq, 32768)) * (r + 1) * 4 * G);
            }
        }
        if (g == 0 && G > 1) {
            rc = rc ? rc : mensure(mm, g, s.oids, nr * 4);
            rc = rc ? rc : mensure(mm, g, s.odists, nr);
            rc = rc ? rc : mensure(mm, g, s.ocounts, static_cast<size_t>(nq) * 4);
        }
        if (rc) return rc;
    }
    MCK(cudaSetDevice(mm->devices[0]));
    MCK(cudaEventRecord(mm->ev[0], mm->ctx[0]->stream));
    for (int g = 0; g < G; ++g) {
        MCK(cudaSetDevice(mm->devices[g]));
        MCK(cudaMemcpyAsync(mm->sh[g].q.p, queries, qbytes, cudaMemcpyHostToDevice, mm->ctx[g]->stream));
    }
    if (ivf && G > 1) {
        // coarse assignment with the cells split over the devices
        for (int g = 0; g < G; ++g) {
            const int base = mm->K / G, extra = mm->K % G;
            const int first = g * base + std::min(g, extra), count = base + (g < extra ? 1 : 0);
            MCTX(g, qadc_coarse_partial_device(mm->ctx[g], mm->sh[g].q.as<float>(), nq, ma, first, count, mm->sh[g].ckeys.as<uint64_t>()));
        }
        std::vector<const void*> send(G);
        std::vector<void*> recv(G);
        for (int g = 0; g < G; ++g) { send[g] = mm->sh[g].ckeys.p; recv[g] = mm->sh[g].gckeys.p; }
        int rc = multi_all_gather(mm, send, recv, static_cast<size_t>(nq) * ma * 8);
        if (rc) return rc;
        for (int g = 0; g < G; ++g)
            MCTX(g, qadc_coarse_merge_device(mm->ctx[g], mm->sh[g].gckeys.as<uint64_t>(), G, nq, ma, mm->sh[g].assign.as<int32_t>()));
    }
    MCK(cudaSetDevice(mm->devices[0]));
    MCK(cudaEventRecord(mm->ev[1], mm->ctx[0]->stream));
    for (int g = 0; g < G; ++g) {
        auto& s = mm->sh[g];
        if (!(ivf && G > 1))
            MCTX(g, qadc_search_device(mm->ctx[g], s.q.as<float>(), nq, ma, r, s.ids.as<uint32_t>(), s.dists.as<int8_t>(),
                                       s.counts.as<int32_t>(), s.keys.as<uint64_t>()));
    }
    if (ivf && G > 1) {
        // "owner computes": every device builds tables and scans keep-prefixes for the probes whose lists it holds, the
        // devices exchange (min entry, r smallest prefix distances) per query, then bounds + int8 tables + list scan
        constexpr int kSub = 32768;
        for (int q0 = 0; q0 < nq; q0 += kSub) {
            const int n = std::min(kSub, nq - q0);
            const size_t lbytes = static_cast<size_t>(n) * (r + 1) * 4;
            std::vector<const void*> send(G);
            std::vector<void*> recv(G);
            for (int g = 0; g < G; ++g) {
                auto& s = mm->sh[g];
                MCTX(g, qadc_tables_local_device(mm->ctx[g], s.q.as<float>() + static_cast<size_t>(q0) * mm->dim,
                                                 s.assign.as<int32_t>() + static_cast<size_t>(q0) * ma, n, ma, r, s.local.as<float>()));
                send[g] = s.local.p; recv[g] = s.glocal.p;
            }
            int rc = multi_all_gather(mm, send, recv, lbytes);
            if (rc) return rc;
            for (int g = 0; g < G; ++g) {
                auto& s = mm->sh[g];
                MCTX(g, qadc_search_bounded_device(mm->ctx[g], s.glocal.as<float>(), G, n, ma, r,
                                                   s.ids.as<uint32_t>() + static_cast<size_t>(q0) * r, s.dists.as<int8_t>() + static_cast<size_t>(q0) * r,
                                                   s.counts.as<int32_t>() + q0, s.keys.as<uint64_t>() + static_cast<size_t>(q0) * r));
            }
            if (!mm->use_nccl && q0 + kSub < nq)   // virtual shards: the share buffers are reused by the next sub-batch
                for (int g = 0; g < G; ++g) MCK(cudaStreamSynchronize(mm->ctx[g]->stream));
        }
    }
    MCK(cudaSetDevice(mm->devices[0]));
    MCK(cudaEventRecord(mm->ev[2], mm->ctx[0]->stream));
    const uint32_t* d_ids = mm->sh[0].ids.as<uint32_t>();
    const int8_t* d_dists = mm->sh[0].dists.as<int8_t>();
    const int32_t* d_counts = mm->sh[0].counts.as<int32_t>();
    if (G > 1) {
        std::vector<GatherJob> jobs(ivf ? 2 : 1);   // labels travel with their keys; a flat id is the key's low 32 bits
        jobs[0].bytes = nr * 8;
        if (ivf) jobs[1].bytes = nr * 4;
        for (int g = 0; g < G; ++g) {
            jobs[0].send.push_back(mm->sh[g].keys.p); jobs[0].recv.push_back(mm->sh[g].gkeys.p);
            if (ivf) { jobs[1].send.push_back(mm->sh[g].ids.p); jobs[1].recv.push_back(mm->sh[g].gids.p); }
        }
        int rc = multi_all_gather(mm, jobs);
        if (rc) return rc;
        MCK(cudaSetDevice(mm->devices[0]));
        MCK(cudaEventRecord(mm->ev[3], mm->ctx[0]->stream));
        auto& s0 = mm->sh[0];
        MCTX(0, qadc_merge_shards_device(mm->ctx[0], s0.gkeys.as<uint64_t>(), ivf ? s0.gids.as<uint32_t>() : nullptr, G, nq, r,
                                         s0.oids.as<uint32_t>(), s0.odists.as<int8_t>(), s0.ocounts.as<int32_t>(), nullptr));
        d_ids = s0.oids.as<uint32_t>(); d_dists = s0.odists.as<int8_t>(); d_counts = s0.ocounts.as<int32_t>();
    } else {
        MCK(cudaEventRecord(mm->ev[3], mm->ctx[0]->stream));
    }
    MCK(cudaSetDevice(mm->devices[0]));
    cudaStream_t st0 = mm->ctx[0]->stream;
    MCK(cudaEventRecord(mm->ev[4], st0));
    MCK(cudaMemcpyAsync(out_ids, d_ids, nr * 4, cudaMemcpyDeviceToHost, st0));
    MCK(cudaMemcpyAsync(out_dists, d_dists, nr, cudaMemcpyDeviceToHost, st0));
    if (out_counts) MCK(cudaMemcpyAsync(out_counts, d_counts, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, st0));
    MCK(cudaEventRecord(mm->ev[5], st0));
    int result = QADC_OK;
    for (int g = 0; g < G; ++g) {
        const int rc = qadc_synchronize(mm->ctx[g]);   // also reports a deferred QADC_EBOUND
        if (rc && result == QADC_OK) result = mfail(mm, rc, std::string("device ") + std::to_string(mm->devices[g]) + ": " + qadc_last_error(mm->ctx[g]));
    }
    if (metrics && result == QADC_OK) {
        MCK(cudaSetDevice(mm->devices[0]));
        float ms = 0;
        auto el = [&](int a, int b) { cudaEventElapsedTime(&ms, mm->ev[a], mm->ev[b]); return static_cast<double>(ms) * 1e3; };
        // device 0's view of the batch; the library's own phase split (tables vs scan) of shard 0
        qadc_ctx* c0 = mm->ctx[0];
        float t_idx = 0, t_tab = 0, t_scan = 0;
        cudaEventElapsedTime(&t_idx, c0->ev[0], c0->ev[1]);
        cudaEventElapsedTime(&t_tab, c0->ev[1], c0->ev[2]);
        cudaEventElapsedTime(&t_scan, c0->ev[2], c0->ev[4]);
        metrics->index_us = el(0, 1) + t_idx * 1e3;          // H2D of the queries + coarse assignment (split + all-gather + merge)
        metrics->rotate_us = 0;
        metrics->table_us = t_tab * 1e3;
        metrics->scan_us = t_scan * 1e3 + el(2, 4);          // scan + shard exchange + merge
        metrics->h2d_us = 0;
        metrics->d2h_us = el(4, 5);
    }
    return result;
}

}  // extern "C"
