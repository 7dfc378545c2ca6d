// torch glue for the tensor-core training path (sgns_tile.cu) and its descriptor probes (umma_probe.cu).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_runtime.h>
#include <vector>

#include "launchers.h"
#include "sgns_params.h"
#include "sgns_tile.h"

namespace {

using torch::Tensor;

inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }

void check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    TORCH_CHECK(e == cudaSuccess, what, ": ", cudaGetErrorString(e));
}

// image: uint8 shared-memory image (copied to a 1024-byte aligned base); ops: int64 [n, 3] =
// {A descriptor, B descriptor (start addresses relative to the image), idesc | accumulate << 32 | tmem column << 40}
Tensor umma_probe(Tensor image, Tensor ops, int64_t ncols) {
    TORCH_CHECK(image.is_cuda() && ops.is_cuda() && image.is_contiguous() && ops.is_contiguous(), "cuda contiguous inputs");
    TORCH_CHECK(image.scalar_type() == torch::kUInt8 && ops.scalar_type() == torch::kInt64 && ops.size(1) == 3, "dtypes");
    c10::cuda::CUDAGuard guard(image.device());
    auto out = torch::zeros({128, ncols}, image.options().dtype(torch::kFloat32));
    int rc = gw2v::launch_umma_probe(image.data_ptr<uint8_t>(), (int)image.numel(),
                                     reinterpret_cast<const unsigned long long*>(ops.data_ptr<int64_t>()),
                                     (int)ops.size(0), out.data_ptr<float>(), (int)ncols, cur_stream());
    TORCH_CHECK(rc == 0, "umma_probe: bad arguments");
    check_launch("umma_probe");
    return out;
}

// table [R, C] fp32; rows int32 [4 * n4]; returns the n4 * 512 shared-memory bytes written by n4 gather4 copies
Tensor gather4_probe(Tensor table, Tensor rows, int64_t col, int64_t box_cols, int64_t bytes_per_op, int64_t swizzle32) {
    TORCH_CHECK(table.is_cuda() && rows.is_cuda() && table.is_contiguous() && rows.is_contiguous(), "cuda contiguous inputs");
    TORCH_CHECK(table.scalar_type() == torch::kFloat32 && rows.scalar_type() == torch::kInt32 && rows.numel() % 4 == 0, "dtypes");
    c10::cuda::CUDAGuard guard(table.device());
    const int n4 = (int)(rows.numel() / 4);
    auto out = torch::zeros({(int64_t)n4 * 512}, table.options().dtype(torch::kUInt8));
    int rc = gw2v::launch_gather4_probe(table.data_ptr<float>(), table.size(0), (int)table.size(1), rows.data_ptr<int>(),
                                        n4, (int)col, (int)box_cols, (int)bytes_per_op, (int)swizzle32, out.data_ptr<uint8_t>(), cur_stream());
    TORCH_CHECK(rc == 0, "gather4_probe failed (rc=", rc, "): 1 = bad size, 2 = tensor map encode failed");
    check_launch("gather4_probe");
    return out;
}

// One tensor-core step: window masks + pair count (pairgen.cu), shared negatives of the tiles, tile kernel.
void sgns_step_tile(Tensor syn0, Tensor syn1, Tensor tokens, Tensor sent_id, Tensor n_tokens, int64_t max_tokens,
                    Tensor alias, Tensor stats, int64_t pos0, int64_t seed, int64_t iteration, int64_t window,
                    int64_t negatives, int64_t window_mode, double alpha, double max_grad, bool compute_loss,
                    int64_t grid, int64_t debug, Tensor cinfo, Tensor pair_off, Tensor n_pairs, Tensor tile_ws,
                    Tensor tile_negs, int64_t tile_negatives, c10::optional<Tensor> exp_table,
                    c10::optional<Tensor> row_scale0, c10::optional<Tensor> row_scale1, c10::optional<Tensor> dbg,
                    int64_t world, int64_t rank, std::vector<int64_t> xbuf_ptrs, std::vector<int64_t> flag_ptrs,
                    c10::optional<Tensor> cta_seq, c10::optional<Tensor> error_flag, c10::optional<Tensor> timing,
                    double tile_neg_scale, double tile_neg_weight) {
    TORCH_CHECK(syn0.is_cuda() && syn1.is_cuda() && syn0.is_contiguous() && syn1.is_contiguous(), "syn0/syn1: cuda contiguous");
    TORCH_CHECK(syn0.scalar_type() == torch::kFloat32 && syn1.scalar_type() == torch::kFloat32, "syn0/syn1 must be fp32");
    TORCH_CHECK(tokens.scalar_type() == torch::kInt32 && sent_id.scalar_type() == torch::kInt32 &&
                n_tokens.scalar_type() == torch::kInt32 && alias.scalar_type() == torch::kInt32 &&
                cinfo.scalar_type() == torch::kInt32 && pair_off.scalar_type() == torch::kInt32 &&
                n_pairs.scalar_type() == torch::kInt32 && tile_ws.scalar_type() == torch::kInt32 &&
                tile_negs.scalar_type() == torch::kInt32 && stats.scalar_type() == torch::kFloat32, "dtypes");
    const int K = (int)syn0.size(1);
    TORCH_CHECK(gw2v::sgns_tile_supported(K, (int)window, (int)negatives, 128, (int)tile_negatives),
                "sgns_tile: unsupported shape (K % 4 == 0, window <= 11, tile_centres = 128, tile_negatives in {32, 64})");
    TORCH_CHECK(max_tokens <= gw2v::pairgen_max_tokens(), "step too large (max ", gw2v::pairgen_max_tokens(), " tokens)");
    TORCH_CHECK(cinfo.numel() >= max_tokens && pair_off.numel() >= max_tokens, "workspaces too small");
    TORCH_CHECK(tile_ws.numel() >= gw2v::pairgen_max_blocks((int)max_tokens), "tile workspace too small");
    TORCH_CHECK(tile_negs.numel() >= (int64_t)gw2v::sgns_tile_max_tiles((int)max_tokens) * tile_negatives, "tile_negs too small");
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
    p.K = K;
    p.window = (int)window; p.negatives = (int)negatives; p.window_mode = (int)window_mode;
    p.alpha = (float)alpha; p.max_grad = (float)max_grad; p.compute_loss = compute_loss ? 1 : 0;
    p.exp_table = nullptr;
    if (exp_table.has_value()) {
        TORCH_CHECK(exp_table->is_cuda() && exp_table->numel() == 1000 && exp_table->scalar_type() == torch::kFloat32, "exp_table");
        p.exp_table = exp_table->data_ptr<float>();
    }
    p.debug = (int)debug;
    p.world = (int)world; p.rank = (int)rank;
    if (world > 1) {
        // column shards: partial dot products are exchanged inside the kernel through peer-mapped symmetric memory
        TORCH_CHECK(world <= gw2v::MAX_WORLD, "world size > ", gw2v::MAX_WORLD, " not supported");
        TORCH_CHECK((int64_t)xbuf_ptrs.size() == world && (int64_t)flag_ptrs.size() == world, "peer pointer lists");
        TORCH_CHECK(cta_seq.has_value() && error_flag.has_value(), "cta_seq / error_flag required");
        TORCH_CHECK(cta_seq->scalar_type() == torch::kInt32 && cta_seq->numel() >= grid, "cta_seq");
        for (int r = 0; r < world; ++r) {
            p.xbuf[r] = reinterpret_cast<float*>(xbuf_ptrs[r]);
            p.flags[r] = reinterpret_cast<uint32_t*>(flag_ptrs[r]);
        }
        p.cta_seq = reinterpret_cast<uint32_t*>(cta_seq->data_ptr<int>());
        p.error_flag = error_flag->data_ptr<int>();
        p.timing = timing.has_value() ? reinterpret_cast<unsigned long long*>(timing->data_ptr<int64_t>()) : nullptr;
    }
    auto opt_f = [&](c10::optional<Tensor>& t, int64_t n, const char* what) -> float* {
        if (!t.has_value()) return nullptr;
        TORCH_CHECK(t->is_cuda() && t->is_contiguous() && t->scalar_type() == torch::kFloat32 && t->numel() >= n, what);
        return t->data_ptr<float>();
    };
    p.tile_neg_scale = (float)tile_neg_scale;
    p.tile_neg_weight = (float)tile_neg_weight;
    gw2v::TileLaunch l{};
    l.cinfo = reinterpret_cast<const uint32_t*>(cinfo.data_ptr<int>());
    l.tile_negs = tile_negs.data_ptr<int>();
    l.n_pairs = n_pairs.data_ptr<int>();
    TORCH_CHECK(row_scale0.has_value() == row_scale1.has_value(), "row_scale0 / row_scale1 come together");
    if (row_scale0.has_value()) {
        TORCH_CHECK(row_scale0->numel() == row_scale1->numel(), "row scale tables must have the same length");
        p.hot_rows = (int)std::min<int64_t>(row_scale0->numel(), syn0.size(0));
        p.row_scale0 = opt_f(row_scale0, 0, "row_scale0");
        p.row_scale1 = opt_f(row_scale1, 0, "row_scale1");
    }
    l.dbg = opt_f(dbg, 128 * (160 + tile_negatives), "dbg");
    l.max_tokens = (int)max_tokens;
    l.tile_negatives = (int)tile_negatives;
    l.grid = (int)grid;
    gw2v::launch_paircount(p.tokens, p.sent_id, p.n_tokens, (int)max_tokens, p.seed_lo, p.seed_hi, p.iteration, p.pos0,
                           p.window, p.window_mode, reinterpret_cast<uint32_t*>(cinfo.data_ptr<int>()),
                           pair_off.data_ptr<int>(), n_pairs.data_ptr<int>(), tile_ws.data_ptr<int>(),
                           stats.data_ptr<float>(), cur_stream());
    int rc = gw2v::launch_sgns_tile(p, l, cur_stream());
    TORCH_CHECK(rc == 0, "sgns_tile launch failed (rc=", rc, "): 1 = unsupported, 2 = tensor map encode failed");
    check_launch("sgns_step_tile");
}

}  // namespace

void register_tile_bindings(py::module& m) {
    m.def("umma_probe", &umma_probe);
    m.def("gather4_probe", &gather4_probe);
    m.def("sgns_step_tile", &sgns_step_tile);
    m.def("sgns_tile_supported", [](int64_t K, int64_t w, int64_t n, int64_t tc, int64_t tn) {
        return gw2v::sgns_tile_supported((int)K, (int)w, (int)n, (int)tc, (int)tn); });
    m.def("sgns_tile_max_tiles", [](int64_t t) { return (int64_t)gw2v::sgns_tile_max_tiles((int)t); });
    m.def("sgns_tile_exchange_geometry", [](int64_t window, int64_t window_mode, int64_t tn) {
        int slots, fl; gw2v::sgns_tile_exchange_geometry((int)window, (int)window_mode, (int)tn, &slots, &fl);
        return std::vector<int64_t>{slots, fl}; });
}
