// torch <-> kernel glue.  Only this file sees torch headers; the kernels take
// raw pointers + a stream (csrc/launchers.h, csrc/sgns_params.h).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_runtime.h>
#include <vector>

#include "launchers.h"
#include "sgns_params.h"
#include "nn_tc.h"
#include "serve_params.h"

namespace {

using torch::Tensor;

#define CHECK_CUDA(x) TORCH_CHECK((x).is_cuda(), #x " must be a CUDA tensor")
#define CHECK_CONTIG(x) TORCH_CHECK((x).is_contiguous(), #x " must be contiguous")
#define CHECK_DT(x, dt) TORCH_CHECK((x).scalar_type() == dt, #x " has wrong dtype")

inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }

void check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    TORCH_CHECK(e == cudaSuccess, what, ": ", cudaGetErrorString(e));
}

// pair generation + the pair-parallel training kernel (single shard or column shards with in-kernel exchange)
void sgns_step_pairs(Tensor syn0, Tensor syn1, Tensor tokens, Tensor sent_id, Tensor n_tokens, int64_t max_tokens,
                     Tensor alias, Tensor stats, int64_t pos0, int64_t seed, int64_t iteration, int64_t window,
                     int64_t negatives, int64_t window_mode, double alpha, double max_grad, bool compute_loss,
                     int64_t grid, int64_t world, int64_t rank, std::vector<int64_t> xbuf_ptrs,
                     std::vector<int64_t> flag_ptrs, c10::optional<Tensor> warp_seq, c10::optional<Tensor> error_flag,
                     c10::optional<Tensor> timing, int64_t debug, Tensor cinfo, Tensor pair_off, Tensor n_pairs,
                     Tensor desc, Tensor tile_ws, int64_t xbuf_mc, int64_t share_centre,
                     c10::optional<Tensor> exp_table, c10::optional<Tensor> row_scale0,
                     c10::optional<Tensor> row_scale1) {
    CHECK_CUDA(syn0); CHECK_CUDA(syn1); CHECK_CONTIG(syn0); CHECK_CONTIG(syn1);
    CHECK_DT(syn0, torch::kFloat32); CHECK_DT(syn1, torch::kFloat32);
    CHECK_DT(tokens, torch::kInt32); CHECK_DT(sent_id, torch::kInt32); CHECK_DT(n_tokens, torch::kInt32);
    CHECK_DT(alias, torch::kInt32); CHECK_DT(stats, torch::kFloat32); CHECK_DT(cinfo, torch::kInt32);
    CHECK_DT(pair_off, torch::kInt32); CHECK_DT(n_pairs, torch::kInt32); CHECK_DT(desc, torch::kInt32);
    CHECK_DT(tile_ws, torch::kInt32);
    TORCH_CHECK(gw2v::sgns_pairs_supported((int)syn0.size(1), (int)window, (int)negatives), "sgns_pairs: unsupported shape");
    const int pd = gw2v::pairgen_desc_ints((int)negatives);
    TORCH_CHECK(cinfo.numel() >= max_tokens && pair_off.numel() >= max_tokens, "pairgen workspaces too small");
    TORCH_CHECK(desc.numel() >= max_tokens * 2 * window * pd * gw2v::pairgen_splits((int)negatives), "descriptor buffer too small");
    TORCH_CHECK(tile_ws.numel() >= gw2v::pairgen_max_blocks((int)max_tokens), "tile workspace too small");
    TORCH_CHECK(max_tokens <= gw2v::pairgen_max_tokens(), "step too large for the pair generator (max ", gw2v::pairgen_max_tokens(), " tokens)");
    c10::cuda::CUDAGuard guard(syn0.device());
    gw2v::SgnsParams p{};
    p.syn0 = syn0.data_ptr<float>();
    p.syn1 = syn1.data_ptr<float>();
    p.tokens = tokens.data_ptr<int>();
    p.sent_id = sent_id.data_ptr<int>();
    p.n_tokens = n_tokens.data_ptr<int>();
    p.alias = reinterpret_cast<const int2*>(alias.data_ptr<int>());
    p.stats = stats.data_ptr<float>();
    p.pos0 = (unsigned long long)pos0;
    p.seed_lo = (uint32_t)((uint64_t)seed & 0xFFFFFFFFull);
    p.seed_hi = (uint32_t)(((uint64_t)seed >> 32) & 0xFFFFFFFFull);
    p.iteration = (uint32_t)iteration;
    p.vocab = (int)syn0.size(0);
    p.K = (int)syn0.size(1);
    p.window = (int)window; p.negatives = (int)negatives; p.window_mode = (int)window_mode;
    p.alpha = (float)alpha; p.max_grad = (float)max_grad; p.compute_loss = compute_loss ? 1 : 0;
    p.exp_table = nullptr;
    if (exp_table.has_value()) {
        CHECK_CUDA(*exp_table); CHECK_CONTIG(*exp_table); CHECK_DT(*exp_table, torch::kFloat32);
        TORCH_CHECK(exp_table->numel() == 1000, "exp_table must have 1000 entries");
        p.exp_table = exp_table->data_ptr<float>();
    }
    p.debug = (int)debug;
    p.world = (int)world; p.rank = (int)rank;
    if (row_scale0.has_value() && row_scale1.has_value()) {        // hot-row damping tables (first hot_rows rows)
        CHECK_CUDA(*row_scale0); CHECK_CUDA(*row_scale1); CHECK_DT(*row_scale0, torch::kFloat32);
        CHECK_DT(*row_scale1, torch::kFloat32); CHECK_CONTIG(*row_scale0); CHECK_CONTIG(*row_scale1);
        TORCH_CHECK(row_scale0->numel() == row_scale1->numel(), "row scale tables must have the same length");
        p.hot_rows = (int)std::min<int64_t>(row_scale0->numel(), syn0.size(0));
        p.row_scale0 = row_scale0->data_ptr<float>();
        p.row_scale1 = row_scale1->data_ptr<float>();
    }
    gw2v::launch_pairgen(p.tokens, p.sent_id, p.n_tokens, (int)max_tokens, p.alias, p.vocab, p.seed_lo, p.seed_hi,
                         p.iteration, p.pos0, p.window, p.window_mode, p.negatives, (int)share_centre,
                         reinterpret_cast<uint32_t*>(cinfo.data_ptr<int>()), pair_off.data_ptr<int>(),
                         n_pairs.data_ptr<int>(), desc.data_ptr<int>(),
                         tile_ws.data_ptr<int>(), stats.data_ptr<float>(), cur_stream());
    if (world > 1) {
        TORCH_CHECK(world <= gw2v::MAX_WORLD, "world size > 8 not supported");
        TORCH_CHECK((int64_t)xbuf_ptrs.size() == world && (int64_t)flag_ptrs.size() == world, "peer pointer lists");
        TORCH_CHECK(warp_seq.has_value() && error_flag.has_value(), "warp_seq/error_flag required");
        for (int r = 0; r < world; ++r) {
            p.xbuf[r] = reinterpret_cast<float*>(xbuf_ptrs[r]);
            p.flags[r] = reinterpret_cast<uint32_t*>(flag_ptrs[r]);
        }
        p.error_flag = error_flag->data_ptr<int>();
        p.timing = timing.has_value() ? reinterpret_cast<unsigned long long*>(timing->data_ptr<int64_t>()) : nullptr;
        p.xbuf_mc = reinterpret_cast<float*>(xbuf_mc);
        gw2v::launch_sgns_pairs_multi(p, desc.data_ptr<int>(), n_pairs.data_ptr<int>(), pd, (int)grid,
                                      reinterpret_cast<uint32_t*>(warp_seq->data_ptr<int>()), cur_stream());
    } else {
        gw2v::launch_sgns_pairs(p, desc.data_ptr<int>(), n_pairs.data_ptr<int>(), pd, (int)grid, cur_stream());
    }
    check_launch("sgns_step_pairs");
}

// hammer the 16-byte chunk pattern of the pair exchange across all peers; returns [torn, observed] (collective launch)
Tensor xchg_selftest(std::vector<int64_t> buf_ptrs, int64_t rank, int64_t iters, Tensor result) {
    CHECK_CUDA(result); CHECK_DT(result, torch::kInt64);
    TORCH_CHECK((int64_t)buf_ptrs.size() <= gw2v::MAX_WORLD && result.numel() >= 2, "xchg_selftest: bad arguments");
    c10::cuda::CUDAGuard guard(result.device());
    gw2v::PeerTest t{};
    t.world = (int)buf_ptrs.size(); t.rank = (int)rank; t.iters = (int)iters;
    for (int r = 0; r < t.world; ++r) t.buf[r] = reinterpret_cast<uint32_t*>(buf_ptrs[r]);
    t.result = reinterpret_cast<unsigned long long*>(result.data_ptr<int64_t>());
    gw2v::launch_xchg_selftest(t, cur_stream());
    check_launch("xchg_selftest");
    return result;
}

void subsample_compact(Tensor tok_in, Tensor sid_in, int64_t T, Tensor keep_thresh, int64_t seed,
                       int64_t iteration, int64_t raw_pos0, Tensor tok_out, Tensor sid_out, Tensor count_out,
                       Tensor tile_ws) {
    CHECK_CUDA(tok_in); CHECK_DT(tok_in, torch::kInt32); CHECK_DT(sid_in, torch::kInt32);
    CHECK_DT(keep_thresh, torch::kInt32); CHECK_DT(tile_ws, torch::kInt32);
    TORCH_CHECK(tile_ws.numel() >= gw2v::subsample_max_blocks((int)T), "tile workspace too small");
    TORCH_CHECK(T <= gw2v::subsample_max_tokens(), "step too large for the sub-sampling scan (max ",
                gw2v::subsample_max_tokens(), " tokens)");
    TORCH_CHECK(tok_out.numel() >= T && sid_out.numel() >= T, "output buffers too small");
    c10::cuda::CUDAGuard guard(tok_in.device());
    gw2v::launch_subsample_compact(
        tok_in.data_ptr<int>(), sid_in.data_ptr<int>(), (int)T,
        reinterpret_cast<const uint32_t*>(keep_thresh.data_ptr<int>()),
        (uint32_t)((uint64_t)seed & 0xFFFFFFFFull), (uint32_t)(((uint64_t)seed >> 32) & 0xFFFFFFFFull),
        (uint32_t)iteration, (unsigned long long)raw_pos0, tok_out.data_ptr<int>(), sid_out.data_ptr<int>(),
        count_out.data_ptr<int>(), tile_ws.data_ptr<int>(), cur_stream());
    check_launch("subsample_compact");
}

int64_t subsample_max_blocks(int64_t max_tokens) { return gw2v::subsample_max_blocks((int)max_tokens); }

void zipf_stream(Tensor alias, int64_t seed, int64_t pos0, Tensor out) {
    CHECK_CUDA(alias); CHECK_DT(alias, torch::kInt32); CHECK_DT(out, torch::kInt32);
    c10::cuda::CUDAGuard guard(alias.device());
    gw2v::launch_zipf_stream(reinterpret_cast<const int2*>(alias.data_ptr<int>()), (int)alias.size(0),
                             (uint32_t)((uint64_t)seed & 0xFFFFFFFFull),
                             (uint32_t)(((uint64_t)seed >> 32) & 0xFFFFFFFFull), (unsigned long long)pos0,
                             (int)out.numel(), out.data_ptr<int>(), cur_stream());
    check_launch("zipf_stream");
}

void init_syn0(Tensor syn0, int64_t col_start, int64_t vector_size, int64_t seed) {
    CHECK_CUDA(syn0); CHECK_CONTIG(syn0); CHECK_DT(syn0, torch::kFloat32);
    c10::cuda::CUDAGuard guard(syn0.device());
    gw2v::launch_init_syn0(syn0.data_ptr<float>(), syn0.size(0), (int)syn0.size(1), (int)col_start,
                           (int)vector_size, (uint32_t)((uint64_t)seed & 0xFFFFFFFFull),
                           (uint32_t)(((uint64_t)seed >> 32) & 0xFFFFFFFFull), cur_stream());
    check_launch("init_syn0");
}

Tensor gather_rows(Tensor syn0, Tensor rows) {
    CHECK_CUDA(syn0); CHECK_CUDA(rows); CHECK_DT(rows, torch::kInt64); CHECK_CONTIG(syn0);
    c10::cuda::CUDAGuard guard(syn0.device());
    auto out = torch::empty({rows.numel(), syn0.size(1)}, syn0.options());
    gw2v::launch_gather_rows(syn0.data_ptr<float>(), reinterpret_cast<const long long*>(rows.data_ptr<int64_t>()), (int)rows.numel(), (int)syn0.size(1),
                             out.data_ptr<float>(), cur_stream());
    check_launch("gather_rows");
    return out;
}

Tensor segment_mean_rows(Tensor syn0, Tensor rows, Tensor offsets) {
    CHECK_CUDA(syn0); CHECK_CUDA(rows); CHECK_CUDA(offsets); CHECK_DT(rows, torch::kInt64);
    CHECK_DT(offsets, torch::kInt64);
    c10::cuda::CUDAGuard guard(syn0.device());
    int64_t ns = offsets.numel() - 1;
    auto out = torch::empty({ns, syn0.size(1)}, syn0.options());
    gw2v::launch_segment_mean_rows(syn0.data_ptr<float>(), reinterpret_cast<const long long*>(rows.data_ptr<int64_t>()), reinterpret_cast<const long long*>(offsets.data_ptr<int64_t>()),
                                   (int)ns, (int)syn0.size(1), out.data_ptr<float>(), cur_stream());
    check_launch("segment_mean_rows");
    return out;
}

Tensor row_sqnorm(Tensor syn0) {
    CHECK_CUDA(syn0); CHECK_CONTIG(syn0);
    c10::cuda::CUDAGuard guard(syn0.device());
    auto out = torch::empty({syn0.size(0)}, syn0.options());
    int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
    gw2v::launch_row_sqnorm(syn0.data_ptr<float>(), syn0.size(0), (int)syn0.size(1), out.data_ptr<float>(), sms,
                            cur_stream());
    check_launch("row_sqnorm");
    return out;
}

Tensor scores_rows(Tensor syn0, Tensor qs) {
    CHECK_CUDA(syn0); CHECK_CUDA(qs); CHECK_CONTIG(syn0); CHECK_CONTIG(qs);
    TORCH_CHECK(qs.size(1) == syn0.size(1), "query slice width != shard columns");
    TORCH_CHECK(qs.size(0) * qs.size(1) * 4 <= 200 * 1024, "query batch too large for the CUDA-core path");
    c10::cuda::CUDAGuard guard(syn0.device());
    auto out = torch::empty({qs.size(0), syn0.size(0)}, syn0.options());
    int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
    gw2v::launch_scores_rows(syn0.data_ptr<float>(), syn0.size(0), (int)syn0.size(1), qs.data_ptr<float>(),
                             (int)qs.size(0), out.data_ptr<float>(), sms, cur_stream());
    check_launch("scores_rows");
    return out;
}

// candidate capacity per query of the threshold filter; beyond it the exact chunk-wise path takes over
constexpr int64_t TOPK_FILTER_CAP = 8192;

// Select the candidates of the top-k of `nvalid` rows.  Fast path: threshold from an evenly spread sample + one
// streaming filter pass (serve_fused.cu); if any query overflows the candidate buffer (adversarial data), every
// chunk is ranked exactly instead.  Returns {cand_v [Q, ncand], cand_i [Q, ncand]}.
std::vector<Tensor> topk_candidates(const float* slab, int nsrc, int64_t Q, int64_t vown, int64_t nvalid,
                                    const float* norms_owned, int64_t row_base, int64_t k,
                                    const torch::TensorOptions& fopt) {
    auto iopt = fopt.dtype(torch::kInt64);
    const char* env = getenv("GW2V_TOPK_FILTER");
    const bool use_filter = !(env && env[0] == '0') && k <= 64 && nvalid > TOPK_FILTER_CAP;
    if (use_filter) {
        int ns = gw2v::topk_sample_chunks(nvalid);
        auto cand_s_v = torch::empty({Q, std::max<int64_t>(1, ns * k)}, fopt);
        auto cand_s_i = torch::empty({Q, std::max<int64_t>(1, ns * k)}, iopt);
        auto tau_v = torch::empty({Q, k}, fopt);
        auto tau_i = torch::empty({Q, k}, iopt);
        auto cand_v = torch::empty({Q, TOPK_FILTER_CAP}, fopt);
        auto cand_i = torch::empty({Q, TOPK_FILTER_CAP}, iopt);
        auto counts = torch::empty({Q}, fopt.dtype(torch::kInt32));
        gw2v::launch_topk_select(slab, nsrc, (int)Q, vown, nvalid, norms_owned, row_base, (int)k,
                                 cand_s_v.data_ptr<float>(), reinterpret_cast<long long*>(cand_s_i.data_ptr<int64_t>()),
                                 tau_v.data_ptr<float>(), reinterpret_cast<long long*>(tau_i.data_ptr<int64_t>()),
                                 cand_v.data_ptr<float>(), reinterpret_cast<long long*>(cand_i.data_ptr<int64_t>()),
                                 counts.data_ptr<int>(), (int)TOPK_FILTER_CAP, cur_stream());
        check_launch("topk_select");
        if (counts.max().item<int>() <= TOPK_FILTER_CAP) return {cand_v, cand_i};      // one small D2H sync
    }
    int nchunks = gw2v::topk_owned_num_chunks(nvalid);
    auto cand_v = torch::empty({Q, std::max<int64_t>(1, (int64_t)nchunks * k)}, fopt);
    auto cand_i = torch::empty({Q, std::max<int64_t>(1, (int64_t)nchunks * k)}, iopt);
    gw2v::launch_topk_owned_stage1(slab, nsrc, (int)Q, vown, nvalid, norms_owned, row_base, (int)k,
                                   cand_v.data_ptr<float>(), reinterpret_cast<long long*>(cand_i.data_ptr<int64_t>()),
                                   cur_stream());
    check_launch("topk_stage1");
    if (nchunks == 0) { cand_v.fill_(-3.0e38f); cand_i.fill_(-1); }
    return {cand_v, cand_i};
}

std::vector<Tensor> cosine_topk(Tensor scores, Tensor norms, int64_t k) {
    CHECK_CUDA(scores); CHECK_CUDA(norms); CHECK_CONTIG(scores); CHECK_CONTIG(norms);
    c10::cuda::CUDAGuard guard(scores.device());
    int64_t Q = scores.size(0), V = scores.size(1);
    auto cand = topk_candidates(scores.data_ptr<float>(), 1, Q, V, V, norms.data_ptr<float>(), 0, k, scores.options());
    auto out_v = torch::empty({Q, k}, scores.options());
    auto out_i = torch::empty({Q, k}, scores.options().dtype(torch::kInt64));
    gw2v::launch_topk_merge(cand[0].data_ptr<float>(), reinterpret_cast<const long long*>(cand[1].data_ptr<int64_t>()),
                            (int)cand[0].size(1), (int)Q, (int)k, out_v.data_ptr<float>(),
                            reinterpret_cast<long long*>(out_i.data_ptr<int64_t>()), cur_stream());
    check_launch("cosine_topk");
    return {out_i, out_v};
}

// tcgen05 nearest-neighbour scores: out[q, v] = sum_k bf16/tf32(syn0[v,k]) * q[q,k]
Tensor scores_tc(Tensor syn0, Tensor qs) {
    CHECK_CUDA(syn0); CHECK_CUDA(qs); CHECK_CONTIG(syn0); CHECK_CONTIG(qs);
    c10::cuda::CUDAGuard guard(syn0.device());
    TORCH_CHECK(qs.size(1) == syn0.size(1), "query slice width != shard columns");
    TORCH_CHECK(gw2v::scores_tc_supported((int)syn0.size(1), (int)qs.size(0)), "scores_tc: unsupported shape");
    auto out = torch::empty({qs.size(0), syn0.size(0)}, syn0.options());
    int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
    int qp = gw2v::scores_tc_padded_queries((int)qs.size(0));
    auto qpad = torch::zeros({qp, qs.size(1)}, qs.options());
    qpad.narrow(0, 0, qs.size(0)).copy_(qs);
    int rc = gw2v::launch_scores_tc(syn0.data_ptr<float>(), syn0.size(0), (int)syn0.size(1), qpad.data_ptr<float>(),
                                    (int)qs.size(0), out.data_ptr<float>(), sms, cur_stream());
    TORCH_CHECK(rc == 0, "scores_tc failed (rc=", rc, "): 1 = unsupported shape, 2 = TMA descriptor encode failed");
    check_launch("scores_tc");
    return out;
}

bool scores_tc_supported(int64_t K, int64_t Q) { return gw2v::scores_tc_supported((int)K, (int)Q); }

// ---------------------------------------------------------------------------------------------------------
// Serving over column shards with the collective fused into the kernels (serve_fused.cu / nn_tc.cu).
// One ServeCtx per engine: symmetric flag array pointers, the CTA arrival counter and the running sequence
// number (every rank issues the same operations in the same order, so the numbers agree without traffic).
struct ServeCtx {
    int64_t world, rank;
    std::vector<int64_t> flag_ptrs;
    Tensor done, err;
    int64_t seq = 0;

    ServeCtx(int64_t world_, int64_t rank_, std::vector<int64_t> flag_ptrs_, Tensor done_, Tensor err_)
        : world(world_), rank(rank_), flag_ptrs(std::move(flag_ptrs_)), done(done_), err(err_) {
        TORCH_CHECK(world >= 1 && world <= gw2v::MAX_WORLD, "world size must be 1..", gw2v::MAX_WORLD);
        TORCH_CHECK((int64_t)flag_ptrs.size() == world, "need one flag pointer per rank");
        CHECK_CUDA(done); CHECK_CUDA(err); CHECK_DT(done, torch::kInt32); CHECK_DT(err, torch::kInt32);
    }
    gw2v::ServeSync next() {
        ++seq;
        gw2v::ServeSync s{};
        for (int64_t r = 0; r < world; ++r) s.flags[r] = reinterpret_cast<uint32_t*>(flag_ptrs[r]);
        s.done = reinterpret_cast<unsigned int*>(done.data_ptr<int>());
        s.error_flag = err.data_ptr<int>();
        s.world = (int)world; s.rank = (int)rank; s.seq = (uint32_t)seq;
        return s;
    }
    void wait(const gw2v::ServeSync& s) {
        gw2v::launch_serve_wait(s.flags[rank], (int)world, s.seq, s.error_flag, cur_stream());
        check_launch("serve_wait");
    }
    void barrier() {
        c10::cuda::CUDAGuard guard(done.device());
        gw2v::ServeSync s = next();
        gw2v::launch_serve_barrier(s, cur_stream());
        check_launch("serve_barrier");
    }
};

gw2v::PeerPtrs peer_ptrs(const ServeCtx& c, const std::vector<int64_t>& ptrs) {
    TORCH_CHECK((int64_t)ptrs.size() == c.world, "need one buffer pointer per rank");
    gw2v::PeerPtrs p{};
    for (int64_t r = 0; r < c.world; ++r) p.p[r] = reinterpret_cast<float*>(ptrs[r]);
    return p;
}

// pull(rows): every rank ends up with the full [R, world*K] rows in its own `out` buffer
void serve_gather_push(ServeCtx& c, Tensor syn0, Tensor rows, std::vector<int64_t> out_ptrs, int64_t ldo) {
    CHECK_CUDA(syn0); CHECK_CUDA(rows); CHECK_DT(rows, torch::kInt64); CHECK_CONTIG(syn0); CHECK_CONTIG(rows);
    c10::cuda::CUDAGuard guard(syn0.device());
    gw2v::ServeSync s = c.next();
    gw2v::launch_gather_rows_push(syn0.data_ptr<float>(), reinterpret_cast<const long long*>(rows.data_ptr<int64_t>()),
                                  (int)rows.numel(), (int)syn0.size(1), peer_ptrs(c, out_ptrs), (int)ldo, s,
                                  cur_stream());
    check_launch("gather_rows_push");
    c.wait(s);
}

void serve_segment_mean_push(ServeCtx& c, Tensor syn0, Tensor rows, Tensor offsets, std::vector<int64_t> out_ptrs,
                             int64_t ldo) {
    CHECK_CUDA(syn0); CHECK_CUDA(rows); CHECK_CUDA(offsets); CHECK_DT(rows, torch::kInt64);
    CHECK_DT(offsets, torch::kInt64); CHECK_CONTIG(syn0); CHECK_CONTIG(rows); CHECK_CONTIG(offsets);
    c10::cuda::CUDAGuard guard(syn0.device());
    gw2v::ServeSync s = c.next();
    gw2v::launch_segment_mean_push(syn0.data_ptr<float>(), reinterpret_cast<const long long*>(rows.data_ptr<int64_t>()),
                                   reinterpret_cast<const long long*>(offsets.data_ptr<int64_t>()),
                                   (int)(offsets.numel() - 1), (int)syn0.size(1), peer_ptrs(c, out_ptrs), (int)ldo, s,
                                   cur_stream());
    check_launch("segment_mean_push");
    c.wait(s);
}

// norms(), step 1: partial sums of squares -> the owner's slab [src][vown]
void serve_sqnorm_push(ServeCtx& c, Tensor syn0, std::vector<int64_t> slab_ptrs, int64_t vown) {
    CHECK_CUDA(syn0); CHECK_CONTIG(syn0);
    c10::cuda::CUDAGuard guard(syn0.device());
    gw2v::ServeSync s = c.next();
    int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
    gw2v::launch_row_sqnorm_push(syn0.data_ptr<float>(), syn0.size(0), (int)syn0.size(1), peer_ptrs(c, slab_ptrs),
                                 vown, s, sms, cur_stream());
    check_launch("row_sqnorm_push");
    c.wait(s);
}

// owner sums the S partial slices (optionally sqrt) and all-gathers its slice into every rank's full vector
void serve_reduce_finish_push(ServeCtx& c, Tensor slab_local, int64_t nsrc, int64_t vown, int64_t nvalid,
                              bool take_sqrt, std::vector<int64_t> full_ptrs) {
    CHECK_CUDA(slab_local); CHECK_DT(slab_local, torch::kFloat32);
    c10::cuda::CUDAGuard guard(slab_local.device());
    gw2v::ServeSync s = c.next();
    gw2v::launch_reduce_finish_push(slab_local.data_ptr<float>(), (int)nsrc, vown, nvalid, take_sqrt ? 1 : 0,
                                    peer_ptrs(c, full_ptrs), s, cur_stream());
    check_launch("reduce_finish_push");
    c.wait(s);
}

// partial score tiles -> the owner's slab [src][Q][vown]  (tcgen05 GEMM epilogue or the CUDA-core kernel)
void serve_scores_push(ServeCtx& c, Tensor syn0, Tensor qs, std::vector<int64_t> slab_ptrs, int64_t vown,
                       bool use_tc) {
    CHECK_CUDA(syn0); CHECK_CUDA(qs); CHECK_CONTIG(syn0); CHECK_CONTIG(qs);
    TORCH_CHECK(qs.size(1) == syn0.size(1), "query slice width != shard columns");
    c10::cuda::CUDAGuard guard(syn0.device());
    gw2v::ServeSync s = c.next();
    int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
    if (use_tc) {
        TORCH_CHECK(gw2v::scores_tc_supported((int)syn0.size(1), (int)qs.size(0)), "scores_tc: unsupported shape");
        int qp = gw2v::scores_tc_padded_queries((int)qs.size(0));
        auto qpad = torch::zeros({qp, qs.size(1)}, qs.options());
        qpad.narrow(0, 0, qs.size(0)).copy_(qs);
        int rc = gw2v::launch_scores_tc_push(syn0.data_ptr<float>(), syn0.size(0), (int)syn0.size(1),
                                             qpad.data_ptr<float>(), (int)qs.size(0), peer_ptrs(c, slab_ptrs), vown, s,
                                             sms, cur_stream());
        TORCH_CHECK(rc == 0, "scores_tc_push failed (rc=", rc, ")");
    } else {
        TORCH_CHECK(qs.size(0) * qs.size(1) * 4 <= 200 * 1024, "query batch too large for the CUDA-core path");
        gw2v::launch_scores_rows_push(syn0.data_ptr<float>(), syn0.size(0), (int)syn0.size(1), qs.data_ptr<float>(),
                                      (int)qs.size(0), peer_ptrs(c, slab_ptrs), vown, s, sms, cur_stream());
    }
    check_launch("scores_push");
    c.wait(s);
}

// owner: sum the S partial score slices, / norm, top-k of the owned rows; winners pushed to every rank [q][src][k]
void serve_topk_owned_push(ServeCtx& c, Tensor slab_local, int64_t nsrc, int64_t Q, int64_t vown, int64_t nvalid,
                           Tensor norms_owned, int64_t row_base, int64_t k, std::vector<int64_t> candv_ptrs,
                           std::vector<int64_t> candi_ptrs) {
    CHECK_CUDA(slab_local); CHECK_CUDA(norms_owned);
    c10::cuda::CUDAGuard guard(slab_local.device());
    auto cand = topk_candidates(slab_local.data_ptr<float>(), (int)nsrc, Q, vown, nvalid, norms_owned.data_ptr<float>(),
                                row_base, k, slab_local.options());
    gw2v::ServeSync s = c.next();
    gw2v::PeerIdx pi{};
    TORCH_CHECK((int64_t)candi_ptrs.size() == c.world, "need one candidate-index pointer per rank");
    for (int64_t r = 0; r < c.world; ++r) pi.p[r] = reinterpret_cast<long long*>(candi_ptrs[r]);
    gw2v::launch_topk_merge_push(cand[0].data_ptr<float>(), reinterpret_cast<const long long*>(cand[1].data_ptr<int64_t>()),
                                 (int)cand[0].size(1), (int)Q, (int)k, peer_ptrs(c, candv_ptrs), pi, s, cur_stream());
    check_launch("topk_owned_push");
    c.wait(s);
}

// ---- nearest neighbours with in-epilogue selection (nn_select.cu) --------------------------------------------
Tensor nn_pad_queries(Tensor qs) {
    int qp = gw2v::scores_tc_padded_queries((int)qs.size(0));
    auto qpad = torch::zeros({qp, qs.size(1)}, qs.options());
    qpad.narrow(0, 0, qs.size(0)).copy_(qs);
    return qpad;
}

// cosines of every `stride`-th row: [Q, ceil(rows / stride)] (the threshold sample)
Tensor nn_sample_cosines(Tensor mat, Tensor inv_norm, Tensor qpad, int64_t Q, int64_t stride) {
    CHECK_CUDA(mat); CHECK_CONTIG(mat); CHECK_CUDA(inv_norm); CHECK_CUDA(qpad); CHECK_CONTIG(qpad);
    CHECK_DT(mat, torch::kFloat32); CHECK_DT(inv_norm, torch::kFloat32);
    c10::cuda::CUDAGuard guard(mat.device());
    const int64_t rows = (mat.size(0) + stride - 1) / stride;
    auto out = torch::empty({Q, rows}, mat.options());
    int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
    int rc = gw2v::launch_nn_select(mat.data_ptr<float>(), rows, (int)mat.size(1), stride * mat.size(1),
                                    inv_norm.data_ptr<float>(), stride, qpad.data_ptr<float>(), (int)Q, 1,
                                    out.data_ptr<float>(), rows, nullptr, nullptr, nullptr, 0, sms, cur_stream());
    TORCH_CHECK(rc == 0, "nn_sample_cosines failed (rc=", rc, ")");
    check_launch("nn_sample_cosines");
    return out;
}

// full sweep; rows whose cosine reaches thr[q] are appended to cand[q] -> (cand int32 [Q, cap], count int32 [Q])
std::vector<Tensor> nn_select(Tensor mat, Tensor inv_norm, Tensor qpad, int64_t Q, Tensor thr, int64_t cap) {
    CHECK_CUDA(mat); CHECK_CONTIG(mat); CHECK_CUDA(inv_norm); CHECK_CUDA(qpad); CHECK_CONTIG(qpad); CHECK_CUDA(thr);
    CHECK_DT(mat, torch::kFloat32); CHECK_DT(inv_norm, torch::kFloat32); CHECK_DT(thr, torch::kFloat32);
    TORCH_CHECK(thr.numel() >= Q && inv_norm.numel() >= mat.size(0), "nn_select: thr / inv_norm too short");
    TORCH_CHECK(mat.size(0) < (1ll << 31), "nn_select: at most 2^31 rows per shard");
    c10::cuda::CUDAGuard guard(mat.device());
    auto iopt = mat.options().dtype(torch::kInt32);
    auto cand = torch::empty({Q, cap}, iopt);
    auto count = torch::zeros({Q}, iopt);
    int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
    int rc = gw2v::launch_nn_select(mat.data_ptr<float>(), mat.size(0), (int)mat.size(1), mat.size(1),
                                    inv_norm.data_ptr<float>(), 1, qpad.data_ptr<float>(), (int)Q, 0, nullptr, 0,
                                    thr.data_ptr<float>(), cand.data_ptr<int>(), count.data_ptr<int>(), (int)cap, sms,
                                    cur_stream());
    TORCH_CHECK(rc == 0, "nn_select failed (rc=", rc, ")");
    check_launch("nn_select");
    return {cand, count};
}

// exact fp32 cosines of the selected rows -> (values [Q, cap] (-3e38 beyond count), global indices int64 [Q, cap])
std::vector<Tensor> nn_rerank(Tensor mat, Tensor inv_norm, Tensor qpad, int64_t Q, Tensor cand, Tensor count,
                              int64_t row_base) {
    CHECK_CUDA(mat); CHECK_CUDA(cand); CHECK_CUDA(count); CHECK_DT(cand, torch::kInt32); CHECK_DT(count, torch::kInt32);
    c10::cuda::CUDAGuard guard(mat.device());
    const int64_t cap = cand.size(1);
    auto out_v = torch::empty({Q, cap}, mat.options());
    auto out_i = torch::empty({Q, cap}, mat.options().dtype(torch::kInt64));
    gw2v::launch_nn_rerank(mat.data_ptr<float>(), (int)mat.size(1), inv_norm.data_ptr<float>(), qpad.data_ptr<float>(),
                           (int)Q, cand.data_ptr<int>(), count.data_ptr<int>(), (int)cap, row_base,
                           out_v.data_ptr<float>(), reinterpret_cast<long long*>(out_i.data_ptr<int64_t>()), cur_stream());
    check_launch("nn_rerank");
    return {out_v, out_i};
}

// column shard -> row shards of the serving replica on every owner (in-kernel peer stores + sequence flag)
void serve_rowshard_push(ServeCtx& c, Tensor syn0, std::vector<int64_t> replica_ptrs, int64_t vown, int64_t ldr) {
    CHECK_CUDA(syn0); CHECK_CONTIG(syn0); CHECK_DT(syn0, torch::kFloat32);
    c10::cuda::CUDAGuard guard(syn0.device());
    gw2v::ServeSync s = c.next();
    int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
    gw2v::launch_rowshard_push(syn0.data_ptr<float>(), syn0.size(0), (int)syn0.size(1), peer_ptrs(c, replica_ptrs), vown,
                               (int)ldr, (int)(c.rank * syn0.size(1)), s, sms, cur_stream());
    check_launch("rowshard_push");
    c.wait(s);
}

// local top-k of a candidate list, winners pushed to every rank's [q][src][k] exchange region
void serve_topk_cand_push(ServeCtx& c, Tensor cand_v, Tensor cand_i, int64_t k, std::vector<int64_t> candv_ptrs,
                          std::vector<int64_t> candi_ptrs) {
    CHECK_CUDA(cand_v); CHECK_CUDA(cand_i); CHECK_CONTIG(cand_v); CHECK_CONTIG(cand_i); CHECK_DT(cand_i, torch::kInt64);
    c10::cuda::CUDAGuard guard(cand_v.device());
    gw2v::ServeSync s = c.next();
    gw2v::PeerIdx pi{};
    TORCH_CHECK((int64_t)candi_ptrs.size() == c.world, "need one candidate-index pointer per rank");
    for (int64_t r = 0; r < c.world; ++r) pi.p[r] = reinterpret_cast<long long*>(candi_ptrs[r]);
    gw2v::launch_topk_merge_push(cand_v.data_ptr<float>(), reinterpret_cast<const long long*>(cand_i.data_ptr<int64_t>()),
                                 (int)cand_v.size(1), (int)cand_v.size(0), (int)k, peer_ptrs(c, candv_ptrs), pi, s,
                                 cur_stream());
    check_launch("topk_cand_push");
    c.wait(s);
}

// final merge of the world*k candidates per query (destroys cand_v)
std::vector<Tensor> serve_topk_final(Tensor cand_v, Tensor cand_i, int64_t k) {
    CHECK_CUDA(cand_v); CHECK_CUDA(cand_i); CHECK_CONTIG(cand_v); CHECK_CONTIG(cand_i);
    c10::cuda::CUDAGuard guard(cand_v.device());
    int64_t Q = cand_v.size(0), ncand = cand_v.size(1);
    auto out_v = torch::empty({Q, k}, cand_v.options());
    auto out_i = torch::empty({Q, k}, cand_i.options());
    gw2v::launch_topk_merge(cand_v.data_ptr<float>(), reinterpret_cast<const long long*>(cand_i.data_ptr<int64_t>()),
                            (int)ncand, (int)Q, (int)k, out_v.data_ptr<float>(),
                            reinterpret_cast<long long*>(out_i.data_ptr<int64_t>()), cur_stream());
    check_launch("topk_merge");
    return {out_i, out_v};
}

// tiny all-gather: every rank's n floats -> [world, n] on every rank
void serve_push_block(ServeCtx& c, Tensor src, std::vector<int64_t> dst_ptrs) {
    CHECK_CUDA(src); CHECK_CONTIG(src); CHECK_DT(src, torch::kFloat32);
    c10::cuda::CUDAGuard guard(src.device());
    gw2v::ServeSync s = c.next();
    gw2v::launch_push_block(src.data_ptr<float>(), src.numel(), peer_ptrs(c, dst_ptrs), s, cur_stream());
    check_launch("push_block");
    c.wait(s);
}

}  // namespace

void register_tile_bindings(py::module& m);      // bindings_tile.cpp

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    register_tile_bindings(m);
    py::class_<ServeCtx>(m, "ServeCtx")
        .def(py::init<int64_t, int64_t, std::vector<int64_t>, Tensor, Tensor>())
        .def("barrier", &ServeCtx::barrier)
        .def_readonly("seq", &ServeCtx::seq);
    m.def("serve_gather_push", &serve_gather_push);
    m.def("serve_segment_mean_push", &serve_segment_mean_push);
    m.def("serve_sqnorm_push", &serve_sqnorm_push);
    m.def("serve_reduce_finish_push", &serve_reduce_finish_push);
    m.def("serve_scores_push", &serve_scores_push);
    m.def("serve_topk_owned_push", &serve_topk_owned_push);
    m.def("serve_topk_final", &serve_topk_final);
    m.def("serve_push_block", &serve_push_block);
    m.def("sgns_step_pairs", &sgns_step_pairs);
    m.def("sgns_pairs_supported", [](int64_t K, int64_t w, int64_t n) { return gw2v::sgns_pairs_supported((int)K, (int)w, (int)n); });
    m.def("sgns_pairs_grid", [](int64_t K, int64_t dev, bool multi) {
        c10::cuda::CUDAGuard guard((c10::DeviceIndex)dev);
        return (int64_t)gw2v::sgns_pairs_grid((int)K, (int)dev, multi); });
    m.def("xchg_selftest", &xchg_selftest);
    m.def("sgns_pairs_multi_geometry", []() {
        int w, ns, sf; gw2v::sgns_pairs_multi_geometry(&w, &ns, &sf); return std::vector<int64_t>{w, ns, sf}; });
    m.def("pairgen_max_blocks", [](int64_t t) { return (int64_t)gw2v::pairgen_max_blocks((int)t); });
    m.def("pairgen_desc_ints", [](int64_t n) { return (int64_t)gw2v::pairgen_desc_ints((int)n); });
    m.def("pairgen_splits", [](int64_t n) { return (int64_t)gw2v::pairgen_splits((int)n); });
    m.def("subsample_compact", &subsample_compact);
    m.def("subsample_max_blocks", &subsample_max_blocks);
    m.def("zipf_stream", &zipf_stream);
    m.def("init_syn0", &init_syn0);
    m.def("gather_rows", &gather_rows);
    m.def("segment_mean_rows", &segment_mean_rows);
    m.def("row_sqnorm", &row_sqnorm);
    m.def("scores_rows", &scores_rows);
    m.def("cosine_topk", &cosine_topk);
    m.def("scores_tc", &scores_tc);
    m.def("scores_tc_supported", &scores_tc_supported);
    m.def("nn_select_supported", [](int64_t K, int64_t Q) { return gw2v::nn_select_supported((int)K, (int)Q); });
    m.def("nn_pad_queries", &nn_pad_queries);
    m.def("nn_sample_cosines", &nn_sample_cosines);
    m.def("nn_select", &nn_select);
    m.def("nn_rerank", &nn_rerank);
    m.def("serve_rowshard_push", &serve_rowshard_push);
    m.def("serve_topk_cand_push", &serve_topk_cand_push);
}
