// extern "C" boundary of libnnpops_b200.so; see include/nnpops_b200.h for the contract of every entry point.
#include "../../include/nnpops_b200.h"
#include <algorithm>
#include <cstring>
#include <string>
#include "ani_model.cuh"
#include <cstdlib>

namespace nnpops {
void batched_linear_forward(const float*, const float*, const float*, float*, int, int, int, int, int, cudaStream_t);
void batched_linear_backward(const float*, const float*, float*, int, int, int, int, cudaStream_t);
template <typename T>
void neighbor_pairs(const T*, const T*, int, T, long long, int*, T*, T*, int*, cudaStream_t);
template <typename T>
void neighbor_pairs_backward(const int*, const T*, const T*, const T*, const T*, long long, int, T*, cudaStream_t);
double gemm_tcgen05_bench(int, int, int, int, int, int);
void gemm_tcgen05_set_streaming(bool);
class CFConvNeighborList;
class CFConvFilter;
CFConvNeighborList* cfconv_neighbors_create(int, float);
void cfconv_neighbors_destroy(CFConvNeighborList*);
void cfconv_neighbors_build(CFConvNeighborList*, const float*, const float*, cudaStream_t);
long long cfconv_neighbors_pairs(const CFConvNeighborList*);
CFConvFilter* cfconv_create(int, int, float, float, int, const float*, const float*, const float*, const float*, int);
void cfconv_destroy(CFConvFilter*);
void cfconv_compute(const CFConvFilter*, const CFConvNeighborList*, const float*, float*, cudaStream_t);
void cfconv_backprop(const CFConvFilter*, const CFConvNeighborList*, const float*, const float*, float*, float*, cudaStream_t);
void pme_direct(const float*, const float*, const int*, const float*, const float*, const int*, int, long long, int, float, float, float*,
                float*, float*, cudaStream_t);
void pme_direct_fused(const float*, const float*, const float*, const int*, int, int, float, float, float, int, int, float*, float*, float*,
                      cudaStream_t);
void pme_reciprocal_forward(const float*, const float*, const float*, int, int, int, int, int, float, float, const float*, const float*,
                            const float*, float*, float*, cudaStream_t);
void pme_reciprocal_backward(const float*, const float*, const float*, int, int, int, int, int, float, const float*, float*, float*,
                             cudaStream_t);
void pme_spread(const float*, const float*, const float*, int, int, int, int, int, float, float*, cudaStream_t);
void pme_solve(float*, const float*, int, int, int, float, const float*, const float*, const float*, float*, float*, cudaStream_t);
}

using namespace nnpops;

namespace {
thread_local std::string g_error;

template <typename F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    } catch (...) {
        g_error = "unknown error";
        return 2;
    }
}

void require_device() {
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0) {
        cudaGetLastError();
        throw std::runtime_error("nnpops_b200 needs a CUDA device (there is no CPU fallback)");
    }
}
}  // namespace

struct nnpops_ani {
    AniAev* impl;
};
struct nnpops_ani_model {
    AniModel* impl;
    float* dPos = nullptr;   // staging for the host-buffer entry point
    float* dBox = nullptr;
    float* dGrad = nullptr;
    float* dEnergy = nullptr;
    int n = 0;
    // The host-buffer entry point works on handle-owned device buffers, so its whole kernel sequence (~25 launches on two streams,
    // 48 tensor-map encodes) is captured once into a CUDA graph and replayed: one launch call per evaluation instead of ~25.
    cudaGraphExec_t graphExec = nullptr;
    cudaStream_t capStream = nullptr;
    int graphHasBox = -1;
    int hostCalls = 0;
};

struct nnpops_cfconv_neighbors {
    CFConvNeighborList* impl;
};
struct nnpops_cfconv {
    CFConvFilter* impl;
};

extern "C" {

const char* nnpops_last_error(void) { return g_error.c_str(); }
int nnpops_abi_version(void) { return 1; }

int nnpops_ani_create(nnpops_ani_t* out, int num_atoms, int num_species, float radial_cutoff, float angular_cutoff,
                      const int* atom_species, int n_radial, const float* radial_fn, int n_angular, const float* angular_fn,
                      int torchani, int max_radial_neighbors, int max_angular_neighbors) {
    return guarded([&] {
        require_device();
        NNP_REQUIRE(out != nullptr, "out must not be NULL");
        auto* h = new nnpops_ani;
        h->impl = nullptr;
        try {
            h->impl = new AniAev(num_atoms, num_species, radial_cutoff, angular_cutoff, atom_species, n_radial, radial_fn, n_angular,
                                 angular_fn, torchani != 0, max_radial_neighbors, max_angular_neighbors);
        } catch (...) {
            delete h;
            throw;
        }
        *out = h;
    });
}

void nnpops_ani_destroy(nnpops_ani_t h) {
    if (!h) return;
    delete h->impl;
    delete h;
}

int nnpops_ani_forward(nnpops_ani_t h, const float* positions, const float* box, float* radial, float* angular, void* stream) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl, "invalid handle");
        h->impl->forward(positions, box, radial, h->impl->radialWidth(), angular, h->impl->angularWidth(), (cudaStream_t)stream);
    });
}

int nnpops_ani_backward(nnpops_ani_t h, const float* radial_grad, const float* angular_grad, float* position_grad, void* stream) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl, "invalid handle");
        h->impl->backward(radial_grad, h->impl->radialWidth(), angular_grad, h->impl->angularWidth(), position_grad, (cudaStream_t)stream);
    });
}

int nnpops_ani_overflowed(nnpops_ani_t h, int* flags) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl && flags, "invalid argument");
        *flags = h->impl->overflowed();
    });
}

int nnpops_ani_overflow_poll(nnpops_ani_t h, int* flags, int* max_radial_neighbors, int* max_angular_neighbors) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl, "invalid handle");
        if (flags) *flags = h->impl->overflowPoll();
        if (max_radial_neighbors) *max_radial_neighbors = h->impl->maxRadialNeighbors();
        if (max_angular_neighbors) *max_angular_neighbors = h->impl->maxAngularNeighbors();
    });
}

int nnpops_ani_work(nnpops_ani_t h, long long* triples, long long* radial_pairs, void* stream) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl, "invalid handle");
        if (triples) *triples = h->impl->countTriples((cudaStream_t)stream);
        if (radial_pairs) *radial_pairs = h->impl->countRadialPairs((cudaStream_t)stream);
    });
}

int nnpops_ani_model_create(nnpops_ani_model_t* out, int num_atoms, int num_species, float radial_cutoff, float angular_cutoff,
                            const int* atom_species, int n_radial, const float* radial_fn, int n_angular, const float* angular_fn,
                            int ensemble_size, int num_layers, const int* dims, const float* params, int mlp_impl,
                            int max_radial_neighbors, int max_angular_neighbors) {
    return nnpops_ani_model_create_sharded(out, num_atoms, num_species, radial_cutoff, angular_cutoff, atom_species, n_radial, radial_fn,
                                           n_angular, angular_fn, ensemble_size, num_layers, dims, params, mlp_impl, max_radial_neighbors,
                                           max_angular_neighbors, 0, 1);
}

int nnpops_ani_model_create_sharded(nnpops_ani_model_t* out, int num_atoms, int num_species, float radial_cutoff, float angular_cutoff,
                                    const int* atom_species, int n_radial, const float* radial_fn, int n_angular, const float* angular_fn,
                                    int ensemble_size, int num_layers, const int* dims, const float* params, int mlp_impl,
                                    int max_radial_neighbors, int max_angular_neighbors, int shard_rank, int shard_count) {
    return guarded([&] {
        require_device();
        NNP_REQUIRE(out != nullptr, "out must not be NULL");
        auto* h = new nnpops_ani_model;
        try {
            h->impl = new AniModel(num_atoms, num_species, radial_cutoff, angular_cutoff, atom_species, n_radial, radial_fn, n_angular,
                                   angular_fn, ensemble_size, num_layers, dims, params, max_radial_neighbors, max_angular_neighbors, true,
                                   shard_rank, shard_count);
            h->impl->mlp().setImpl(mlp_impl == 1 ? MlpImpl::Tcgen05 : MlpImpl::Simt);
            h->n = num_atoms;
        } catch (...) {
            delete h;
            throw;
        }
        *out = h;
    });
}

int nnpops_ani_model_create_owned(nnpops_ani_model_t* out, int num_atoms, int num_species, float radial_cutoff, float angular_cutoff,
                                  const int* atom_species, int n_radial, const float* radial_fn, int n_angular, const float* angular_fn,
                                  int ensemble_size, int num_layers, const int* dims, const float* params, int mlp_impl,
                                  int max_radial_neighbors, int max_angular_neighbors, const unsigned char* owned) {
    return guarded([&] {
        require_device();
        NNP_REQUIRE(out != nullptr && owned != nullptr, "out and owned must not be NULL");
        auto* h = new nnpops_ani_model;
        try {
            h->impl = new AniModel(num_atoms, num_species, radial_cutoff, angular_cutoff, atom_species, n_radial, radial_fn, n_angular,
                                   angular_fn, ensemble_size, num_layers, dims, params, max_radial_neighbors, max_angular_neighbors, true, 0, 1,
                                   owned);
            h->impl->mlp().setImpl(mlp_impl == 1 ? MlpImpl::Tcgen05 : MlpImpl::Simt);
            h->n = num_atoms;
        } catch (...) {
            delete h;
            throw;
        }
        *out = h;
    });
}

void nnpops_ani_model_destroy(nnpops_ani_model_t h) {
    if (!h) return;
    if (h->graphExec) cudaGraphExecDestroy(h->graphExec);
    if (h->capStream) cudaStreamDestroy(h->capStream);
    delete h->impl;
    cudaFree(h->dPos); cudaFree(h->dBox); cudaFree(h->dGrad); cudaFree(h->dEnergy);
    delete h;
}

int nnpops_ani_model_energy_grad(nnpops_ani_model_t h, const float* positions, const float* box, float* energy, float* position_grad,
                                 void* stream) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl, "invalid handle");
        h->impl->energyAndGradient(positions, box, energy, position_grad, (cudaStream_t)stream);
    });
}

int nnpops_ani_model_energy_grad_host(nnpops_ani_model_t h, const float* positions_host, const float* box_host, float* energy_host,
                                      float* position_grad_host, void* stream) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl, "invalid handle");
        cudaStream_t s = (cudaStream_t)stream;
        const size_t n = (size_t)(h->n > 0 ? h->n : 1);
        if (!h->dPos) {
            NNP_CUDA_CHECK(cudaMalloc(&h->dPos, sizeof(float) * 3 * n));
            NNP_CUDA_CHECK(cudaMalloc(&h->dBox, sizeof(float) * 9));
            NNP_CUDA_CHECK(cudaMalloc(&h->dGrad, sizeof(float) * 3 * n));
            NNP_CUDA_CHECK(cudaMalloc(&h->dEnergy, sizeof(float)));
        }
        NNP_CUDA_CHECK(cudaMemcpyAsync(h->dPos, positions_host, sizeof(float) * 3 * h->n, cudaMemcpyHostToDevice, s));
        if (box_host) NNP_CUDA_CHECK(cudaMemcpyAsync(h->dBox, box_host, sizeof(float) * 9, cudaMemcpyHostToDevice, s));
        static const bool noGraph = std::getenv("NNPOPS_NO_GRAPH") != nullptr;
        const int hasBox = box_host ? 1 : 0;
        const bool graphable = !noGraph && !h->impl->timingActive() && h->n > 0;
        if (graphable && h->graphExec && h->graphHasBox == hasBox) {
            NNP_CUDA_CHECK(cudaGraphLaunch(h->graphExec, s));
        } else if (graphable && h->hostCalls >= 1) {
            // the first call ran eagerly (one-time attribute / workspace set-up); capture on a private stream (the caller's may be the
            // legacy default stream, which cannot be captured) and replay on the caller's
            if (h->graphExec) { cudaGraphExecDestroy(h->graphExec); h->graphExec = nullptr; }
            if (!h->capStream) NNP_CUDA_CHECK(cudaStreamCreateWithFlags(&h->capStream, cudaStreamNonBlocking));
            cudaGraph_t graph = nullptr;
            NNP_CUDA_CHECK(cudaStreamBeginCapture(h->capStream, cudaStreamCaptureModeThreadLocal));
            try {
                h->impl->energyAndGradient(h->dPos, box_host ? h->dBox : nullptr, h->dEnergy, h->dGrad, h->capStream);
            } catch (...) {
                cudaStreamEndCapture(h->capStream, &graph);
                if (graph) cudaGraphDestroy(graph);
                throw;
            }
            NNP_CUDA_CHECK(cudaStreamEndCapture(h->capStream, &graph));
            NNP_CUDA_CHECK(cudaGraphInstantiate(&h->graphExec, graph, 0));
            cudaGraphDestroy(graph);
            h->graphHasBox = hasBox;
            NNP_CUDA_CHECK(cudaGraphLaunch(h->graphExec, s));
        } else {
            h->impl->energyAndGradient(h->dPos, box_host ? h->dBox : nullptr, h->dEnergy, h->dGrad, s);
        }
        h->hostCalls++;
        NNP_CUDA_CHECK(cudaMemcpyAsync(energy_host, h->dEnergy, sizeof(float), cudaMemcpyDeviceToHost, s));
        NNP_CUDA_CHECK(cudaMemcpyAsync(position_grad_host, h->dGrad, sizeof(float) * 3 * h->n, cudaMemcpyDeviceToHost, s));
        NNP_CUDA_CHECK(cudaStreamSynchronize(s));
    });
}

int nnpops_ani_model_buffers(nnpops_ani_model_t h, float** features, float** feature_grad, int* stride, int* row_of_atom) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl, "invalid handle");
        if (features) *features = h->impl->features();
        if (feature_grad) *feature_grad = h->impl->featureGrad();
        if (stride) *stride = h->impl->featureStride();
        if (row_of_atom) std::memcpy(row_of_atom, h->impl->rowOfAtom().data(), sizeof(int) * h->impl->rowOfAtom().size());
    });
}

int nnpops_ani_model_read_features(nnpops_ani_model_t h, int which, float* out, void* stream) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl && out, "invalid argument");
        h->impl->readFeatures(which, out, (cudaStream_t)stream);
    });
}

int nnpops_ani_model_work(nnpops_ani_model_t h, long long* triples, long long* radial_pairs, double* mlp_flops_forward, void* stream) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl, "invalid handle");
        if (triples) *triples = h->impl->aev().countTriples((cudaStream_t)stream);
        if (radial_pairs) *radial_pairs = h->impl->aev().countRadialPairs((cudaStream_t)stream);
        if (mlp_flops_forward) *mlp_flops_forward = h->impl->denseFlopsForward();
    });
}

int nnpops_ani_set_skin(nnpops_ani_t h, float skin) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl, "invalid handle");
        h->impl->setSkin(skin);
    });
}

int nnpops_ani_model_set_skin(nnpops_ani_model_t h, float skin) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl, "invalid handle");
        if (h->graphExec) { cudaGraphExecDestroy(h->graphExec); h->graphExec = nullptr; }   // the captured step has the old kernels
        h->impl->aev().setSkin(skin);
    });
}

int nnpops_ani_model_skin_stats(nnpops_ani_model_t h, unsigned long long* rebuilds, unsigned long long* reuses) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl, "invalid handle");
        h->impl->aev().skinStats(rebuilds, reuses);
    });
}

int nnpops_ani_model_mlp_fused(nnpops_ani_model_t h, int* fused) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl && fused, "invalid argument");
        *fused = h->impl->mlp().fused() ? 1 : 0;
    });
}

int nnpops_ani_model_info(nnpops_ani_model_t h, int* aev_length, int* active_features, double* mlp_flops_forward_executed) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl, "invalid handle");
        if (aev_length) *aev_length = h->impl->aevLength();
        if (active_features) *active_features = h->impl->activeFeatures();
        if (mlp_flops_forward_executed) *mlp_flops_forward_executed = h->impl->mlp().flopsForward();
    });
}

int nnpops_ani_model_overflow_poll(nnpops_ani_model_t h, int* flags, int* max_radial_neighbors, int* max_angular_neighbors) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl, "invalid handle");
        if (flags) *flags = h->impl->aev().overflowPoll();
        if (max_radial_neighbors) *max_radial_neighbors = h->impl->aev().maxRadialNeighbors();
        if (max_angular_neighbors) *max_angular_neighbors = h->impl->aev().maxAngularNeighbors();
    });
}

int nnpops_ani_model_overflowed(nnpops_ani_model_t h, int* flags) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl && flags, "invalid argument");
        *flags = h->impl->aev().overflowed();
    });
}

int nnpops_ani_model_timing_begin(nnpops_ani_model_t h, int max_steps) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl && max_steps > 0, "invalid argument");
        h->impl->timingBegin(max_steps);
    });
}

int nnpops_ani_model_timing_end(nnpops_ani_model_t h, float* stage_ms, int* steps) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl && stage_ms && steps, "invalid argument");
        *steps = h->impl->timingEnd(stage_ms);
    });
}

int nnpops_launch_count(unsigned long long* count) {
    return guarded([&] {
        NNP_REQUIRE(count, "invalid argument");
        *count = g_launches;
    });
}

int nnpops_batched_linear_forward(const float* vectors, const float* weights, const float* biases, float* out, int num_atoms,
                                  int num_models, int vec_models, int n_out, int n_in, void* stream) {
    return guarded([&] {
        require_device();
        batched_linear_forward(vectors, weights, biases, out, num_atoms, num_models, vec_models, n_out, n_in, (cudaStream_t)stream);
    });
}

int nnpops_batched_linear_backward(const float* grad_out, const float* weights, float* grad_vectors, int num_atoms, int num_models,
                                   int n_out, int n_in, void* stream) {
    return guarded([&] {
        require_device();
        batched_linear_backward(grad_out, weights, grad_vectors, num_atoms, num_models, n_out, n_in, (cudaStream_t)stream);
    });
}

int nnpops_neighbor_pairs_f32(const float* positions, const float* box, int num_atoms, float cutoff, long long max_num_pairs,
                              int* neighbors, float* deltas, float* distances, int* num_found, void* stream) {
    return guarded([&] {
        require_device();
        neighbor_pairs<float>(positions, box, num_atoms, cutoff, max_num_pairs, neighbors, deltas, distances, num_found, (cudaStream_t)stream);
    });
}

int nnpops_neighbor_pairs_f64(const double* positions, const double* box, int num_atoms, double cutoff, long long max_num_pairs,
                              int* neighbors, double* deltas, double* distances, int* num_found, void* stream) {
    return guarded([&] {
        require_device();
        neighbor_pairs<double>(positions, box, num_atoms, cutoff, max_num_pairs, neighbors, deltas, distances, num_found, (cudaStream_t)stream);
    });
}

int nnpops_neighbor_pairs_backward_f32(const int* neighbors, const float* deltas, const float* distances, const float* grad_deltas,
                                       const float* grad_distances, long long num_pairs, int num_atoms, float* grad_positions, void* stream) {
    return guarded([&] {
        require_device();
        neighbor_pairs_backward<float>(neighbors, deltas, distances, grad_deltas, grad_distances, num_pairs, num_atoms, grad_positions,
                                       (cudaStream_t)stream);
    });
}

int nnpops_neighbor_pairs_backward_f64(const int* neighbors, const double* deltas, const double* distances, const double* grad_deltas,
                                       const double* grad_distances, long long num_pairs, int num_atoms, double* grad_positions,
                                       void* stream) {
    return guarded([&] {
        require_device();
        neighbor_pairs_backward<double>(neighbors, deltas, distances, grad_deltas, grad_distances, num_pairs, num_atoms, grad_positions,
                                        (cudaStream_t)stream);
    });
}

int nnpops_pme_direct(const float* positions, const float* charges, const int* neighbors, const float* deltas, const float* distances,
                      const int* exclusions, int num_atoms, long long num_pairs, int max_exclusions, float alpha, float coulomb, float* energy,
                      float* pos_deriv, float* charge_deriv, void* stream) {
    return guarded([&] {
        require_device();
        pme_direct(positions, charges, neighbors, deltas, distances, exclusions, num_atoms, num_pairs, max_exclusions, alpha, coulomb, energy,
                   pos_deriv, charge_deriv, (cudaStream_t)stream);
    });
}

int nnpops_pme_direct_fused(const float* positions, const float* charges, const float* box, const int* exclusions, int num_atoms,
                            int max_exclusions, float cutoff, float alpha, float coulomb, int shard_index, int shard_count, float* energy,
                            float* pos_deriv, float* charge_deriv, void* stream) {
    return guarded([&] {
        require_device();
        pme_direct_fused(positions, charges, box, exclusions, num_atoms, max_exclusions, cutoff, alpha, coulomb, shard_index, shard_count,
                         energy, pos_deriv, charge_deriv, (cudaStream_t)stream);
    });
}

int nnpops_pme_reciprocal_forward(const float* positions, const float* charges, const float* box, int num_atoms, int gridx, int gridy,
                                  int gridz, int order, float alpha, float coulomb, const float* xmoduli, const float* ymoduli,
                                  const float* zmoduli, float* energy, float* recip_grid, void* stream) {
    return guarded([&] {
        require_device();
        pme_reciprocal_forward(positions, charges, box, num_atoms, gridx, gridy, gridz, order, alpha, coulomb, xmoduli, ymoduli, zmoduli,
                               energy, recip_grid, (cudaStream_t)stream);
    });
}

int nnpops_pme_spread(const float* positions, const float* charges, const float* box, int num_atoms, int gridx, int gridy, int gridz,
                      int order, float coulomb, float* real_grid, void* stream) {
    return guarded([&] {
        require_device();
        NNP_REQUIRE(real_grid != nullptr, "real_grid must not be NULL");
        pme_spread(positions, charges, box, num_atoms, gridx, gridy, gridz, order, coulomb, real_grid, (cudaStream_t)stream);
    });
}

int nnpops_pme_solve(float* real_grid, const float* box, int gridx, int gridy, int gridz, float alpha, const float* xmoduli,
                     const float* ymoduli, const float* zmoduli, float* energy, float* recip_grid, void* stream) {
    return guarded([&] {
        require_device();
        NNP_REQUIRE(real_grid != nullptr && recip_grid != nullptr && energy != nullptr, "real_grid, recip_grid and energy must not be NULL");
        pme_solve(real_grid, box, gridx, gridy, gridz, alpha, xmoduli, ymoduli, zmoduli, energy, recip_grid, (cudaStream_t)stream);
    });
}

int nnpops_pme_reciprocal_backward(const float* positions, const float* charges, const float* box, int num_atoms, int gridx, int gridy,
                                   int gridz, int order, float coulomb, const float* recip_grid, float* pos_deriv, float* charge_deriv,
                                   void* stream) {
    return guarded([&] {
        require_device();
        pme_reciprocal_backward(positions, charges, box, num_atoms, gridx, gridy, gridz, order, coulomb, recip_grid, pos_deriv, charge_deriv,
                                (cudaStream_t)stream);
    });
}

int nnpops_cfconv_neighbors_create(nnpops_cfconv_neighbors_t* out, int num_atoms, float cutoff) {
    return guarded([&] {
        require_device();
        NNP_REQUIRE(out != nullptr, "out must not be NULL");
        *out = new nnpops_cfconv_neighbors{cfconv_neighbors_create(num_atoms, cutoff)};
    });
}

void nnpops_cfconv_neighbors_destroy(nnpops_cfconv_neighbors_t h) {
    if (!h) return;
    cfconv_neighbors_destroy(h->impl);
    delete h;
}

int nnpops_cfconv_neighbors_build(nnpops_cfconv_neighbors_t h, const float* positions, const float* box, void* stream) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl, "invalid handle");
        cfconv_neighbors_build(h->impl, positions, box, (cudaStream_t)stream);
    });
}

int nnpops_cfconv_neighbors_num_pairs(nnpops_cfconv_neighbors_t h, long long* num_pairs) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl && num_pairs, "invalid argument");
        *num_pairs = cfconv_neighbors_pairs(h->impl);
    });
}

int nnpops_cfconv_create(nnpops_cfconv_t* out, int width, int num_gaussians, float cutoff, float gaussian_width, int activation,
                         const float* w1, const float* b1, const float* w2, const float* b2, int points_per_sigma) {
    return guarded([&] {
        require_device();
        NNP_REQUIRE(out != nullptr, "out must not be NULL");
        *out = new nnpops_cfconv{cfconv_create(width, num_gaussians, cutoff, gaussian_width, activation, w1, b1, w2, b2, points_per_sigma)};
    });
}

void nnpops_cfconv_destroy(nnpops_cfconv_t h) {
    if (!h) return;
    cfconv_destroy(h->impl);
    delete h;
}

int nnpops_cfconv_compute(nnpops_cfconv_t h, nnpops_cfconv_neighbors_t neighbors, const float* input, float* output, void* stream) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl && neighbors && neighbors->impl, "invalid handle");
        cfconv_compute(h->impl, neighbors->impl, input, output, (cudaStream_t)stream);
    });
}

int nnpops_cfconv_backprop(nnpops_cfconv_t h, nnpops_cfconv_neighbors_t neighbors, const float* input, const float* output_grad,
                           float* input_grad, float* position_grad, void* stream) {
    return guarded([&] {
        NNP_REQUIRE(h && h->impl && neighbors && neighbors->impl, "invalid handle");
        cfconv_backprop(h->impl, neighbors->impl, input, output_grad, input_grad, position_grad, (cudaStream_t)stream);
    });
}

// fp32 FMA throughput of the device, measured: the denominator of the AEV kernels' arithmetic roofline (bench.py reports it next to the
// figure derived from SM count x 128 lanes x 2 x clock).  8 independent FMA chains per thread, 2048 threads per SM.
__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.678f) out[0] = s;   // never true: keeps the chains alive
}

int nnpops_debug_fma_peak(double* tflops) {
    return guarded([&] {
        require_device();
        NNP_REQUIRE(tflops != nullptr, "tflops must not be NULL");
        int dev = 0, sms = 0;
        NNP_CUDA_CHECK(cudaGetDevice(&dev));
        NNP_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        float* out = nullptr;
        NNP_CUDA_CHECK(cudaMalloc(&out, sizeof(float)));
        const int iters = 4096, grid = sms * 8;
        cudaEvent_t e0, e1;
        NNP_CUDA_CHECK(cudaEventCreate(&e0)); NNP_CUDA_CHECK(cudaEventCreate(&e1));
        fma_peak_kernel<<<grid, 256>>>(out, 64, 0.999f, 0.001f);   // warm-up
        double best = 0.0;
        for (int rep = 0; rep < 5; rep++) {
            NNP_CUDA_CHECK(cudaEventRecord(e0));
            fma_peak_kernel<<<grid, 256>>>(out, iters, 0.999f, 0.001f);
            NNP_CUDA_CHECK(cudaEventRecord(e1));
            NNP_CUDA_CHECK(cudaEventSynchronize(e1));
            float ms = 0.0f;
            NNP_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
            const double flops = 2.0 * 8 * 16 * (double)iters * 256.0 * grid;
            best = std::max(best, flops / (ms * 1e-3) / 1e12);
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
        *tflops = best;
    });
}

int nnpops_debug_gemm_bench(int m, int n, int k, int batch, int mode, int streaming, int iters, double* ms) {
    return guarded([&] {
        require_device();
        gemm_tcgen05_set_streaming(streaming != 0);
        *ms = gemm_tcgen05_bench(m, n, k, batch, mode, iters);
        gemm_tcgen05_set_streaming(false);
    });
}

}  // extern "C"
