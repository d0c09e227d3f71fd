/*
 * nnpops_b200 -- C ABI of the B200-native NNPOps hot path (libnnpops_b200.so).
 *
 * Every entry point takes plain pointers and sizes; device pointers are marked "device".  `stream` is a cudaStream_t passed as
 * void* (NULL = the legacy default stream).  All functions return 0 on success and a non-zero code on failure;
 * nnpops_last_error() then returns a message (thread-local).  Nothing in this library falls back to a CPU path: without a CUDA
 * device every compute call fails with an error.
 *
 * Each block cites the reference interface it replaces (paths relative to the reference repository root).
 */
#ifndef NNPOPS_B200_H
#define NNPOPS_B200_H

#if defined(__GNUC__)
#define NNPOPS_API __attribute__((visibility("default")))
#else
#define NNPOPS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

NNPOPS_API const char* nnpops_last_error(void);
/* ABI version of this header; bumped on any signature change */
NNPOPS_API int nnpops_abi_version(void);

/* --------------------------------------------------------------------------------------------------------------------
 * ANI symmetry functions.  Replaces class ANISymmetryFunctions / CudaANISymmetryFunctions
 *   src/ani/ANISymmetryFunctions.h:60-64   (constructor)            -> nnpops_ani_create
 *   src/ani/ANISymmetryFunctions.h:78      (computeSymmetryFunctions) -> nnpops_ani_forward
 *   src/ani/ANISymmetryFunctions.h:92      (backprop)               -> nnpops_ani_backward
 * which the torch Holder drives from src/pytorch/SymmetryFunctions.cpp:124-133,155,172.
 * radial_fn  : n_radial  x {eta, rs}                (struct RadialFunction,  ANISymmetryFunctions.h:29-32)
 * angular_fn : n_angular x {eta, rs, zeta, thetas}  (struct AngularFunction, ANISymmetryFunctions.h:34-39)
 * max_*_neighbors: capacity of the per-atom neighbour rows (0 = defaults 256 / 96); nnpops_ani_overflowed reports a
 * system denser than that (the reference has no such limit because it stores an N x N table).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct nnpops_ani* nnpops_ani_t;

NNPOPS_API int nnpops_ani_create(nnpops_ani_t* out, int num_atoms, int num_species, float radial_cutoff, float angular_cutoff,
                      const int* atom_species /* host [num_atoms] */, int n_radial, const float* radial_fn /* host */,
                      int n_angular, const float* angular_fn /* host */, int torchani, int max_radial_neighbors,
                      int max_angular_neighbors);
NNPOPS_API void nnpops_ani_destroy(nnpops_ani_t h);
/* positions: device float [num_atoms][3]; box: device float [3][3] or NULL (non-periodic);
 * radial: device float [num_atoms][num_species*n_radial]; angular: device float [num_atoms][num_species(num_species+1)/2*n_angular] */
NNPOPS_API int nnpops_ani_forward(nnpops_ani_t h, const float* positions, const float* box, float* radial, float* angular, void* stream);
/* uses the positions and box of the most recent forward (ANISymmetryFunctions.h:83-84); position_grad: device float [num_atoms][3] */
NNPOPS_API int nnpops_ani_backward(nnpops_ani_t h, const float* radial_grad, const float* angular_grad, float* position_grad, void* stream);
/* synchronises the device; *flags != 0 when a neighbour row overflowed (bit 0 radial, bit 1 angular) */
NNPOPS_API int nnpops_ani_overflowed(nnpops_ani_t h, int* flags);
/* synchronises; work counters of the last forward: triples = sum_i n_i(n_i-1)/2, pairs = undirected pairs within the radial cutoff */
NNPOPS_API int nnpops_ani_work(nnpops_ani_t h, long long* triples, long long* radial_pairs, void* stream);

/* --------------------------------------------------------------------------------------------------------------------
 * Fused ANI model: AEV -> per-species ensemble MLP -> energy and dE/dx.  Replaces the module chain
 *   src/pytorch/OptimizedTorchANI.py:45-54 (forward) + autograd backward, i.e.
 *   SymmetryFunctions.cpp:236-255 (AEV fwd/bwd) and BatchedNN.py:90-111 / BatchedNN.cpp:30-42 (BatchedLinear + CELU chain).
 * dims  : host int [num_species][num_layers+1]; layer l of species s maps dims[s][l] -> dims[s][l+1]; dims[s][0] = AEV length,
 *         dims[s][num_layers] = 1.
 * params: host float, for each species, for each ensemble member, for each layer: W (out x in, row-major, as nn.Linear.weight)
 *         then b (out).
 * mlp_impl: 0 = fp32 SIMT GEMM (validation path), 1 = tcgen05 tensor-core GEMM.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct nnpops_ani_model* nnpops_ani_model_t;

NNPOPS_API int nnpops_ani_model_create(nnpops_ani_model_t* out, int num_atoms, int num_species, float radial_cutoff, float angular_cutoff,
                            const int* atom_species, int n_radial, const float* radial_fn, int n_angular, const float* angular_fn,
                            int ensemble_size, int num_layers, const int* dims, const float* params, int mlp_impl,
                            int max_radial_neighbors, int max_angular_neighbors);
NNPOPS_API void nnpops_ani_model_destroy(nnpops_ani_model_t h);
/* energy: device float[1] (sum over atoms of the ensemble-mean atomic energies, no self-energy shift);
 * position_grad: device float [num_atoms][3] = dE/dx */
NNPOPS_API int nnpops_ani_model_energy_grad(nnpops_ani_model_t h, const float* positions, const float* box, float* energy,
                                 float* position_grad, void* stream);
/* same call with HOST buffers (pinned or pageable): copies in, evaluates, copies energy and gradient out, synchronises */
NNPOPS_API int nnpops_ani_model_energy_grad_host(nnpops_ani_model_t h, const float* positions_host, const float* box_host, float* energy_host,
                                      float* position_grad_host, void* stream);
/* introspection for tests and benchmarks: device pointers to the species-sorted AEV matrix and its gradient, the row stride in
 * floats, and (host, num_atoms ints) the row of each atom */
NNPOPS_API int nnpops_ani_model_buffers(nnpops_ani_model_t h, float** features, float** feature_grad, int* stride, int* row_of_atom);
/* copy the AEV matrix (which = 0) or its gradient (which = 1) of the last evaluation into out: device float [num_atoms][aev_length],
 * ATOM order */
NNPOPS_API int nnpops_ani_model_read_features(nnpops_ani_model_t h, int which, float* out, void* stream);
NNPOPS_API int nnpops_ani_model_work(nnpops_ani_model_t h, long long* triples, long long* radial_pairs, double* mlp_flops_forward, void* stream);
NNPOPS_API int nnpops_ani_model_overflowed(nnpops_ani_model_t h, int* flags);
/* CUDA-event timing of the pipeline stages on the launching stream, for benchmarks: after timing_begin the next max_steps
 * evaluations record events; timing_end synchronises and returns the summed milliseconds of the 7 stages
 * {cell list + neighbour rows, radial fwd, angular fwd, MLP fwd, MLP bwd, radial bwd, angular bwd} and the evaluations seen */
NNPOPS_API int nnpops_ani_model_timing_begin(nnpops_ani_model_t h, int max_steps);
NNPOPS_API int nnpops_ani_model_timing_end(nnpops_ani_model_t h, float* stage_ms /* host [7] */, int* steps);
/* number of CUDA kernels this library has launched so far in this process */
NNPOPS_API int nnpops_launch_count(unsigned long long* count);

/* --------------------------------------------------------------------------------------------------------------------
 * BatchedLinear.  Replaces torch op NNPOpsBatchedNN::BatchedLinear (src/pytorch/BatchedNN.cpp:30-50):
 *   out[a][m][o] = sum_i weights[a][m][o][i] * vectors[a][m or 0][i] + biases[a][m][o]
 * vectors: device [num_atoms][vec_models][in] with vec_models in {1, num_models} (1 = broadcast over the ensemble axis);
 * backward: grad_vectors[a][m][i] = sum_o grad_out[a][m][o] * weights[a][m][o][i]   (per member; the caller sums over m when
 * vec_models == 1, as autograd does for the reference op).
 * ------------------------------------------------------------------------------------------------------------------ */
NNPOPS_API int nnpops_batched_linear_forward(const float* vectors, const float* weights, const float* biases, float* out, int num_atoms,
                                  int num_models, int vec_models, int n_out, int n_in, void* stream);
NNPOPS_API int nnpops_batched_linear_backward(const float* grad_out, const float* weights, float* grad_vectors, int num_atoms, int num_models,
                                   int n_out, int n_in, void* stream);

#ifdef __cplusplus
}
#endif
#endif
