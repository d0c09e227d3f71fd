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
 * max_*_neighbors: capacity of the per-atom neighbour rows (0 = defaults 256 / 64); nnpops_ani_overflowed reports a
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
/* never blocks: the same flags as of the last forward whose device work has completed, plus the row capacities in use (any
 * pointer may be NULL).  The reference has no neighbour limit (N x N table, CudaANISymmetryFunctions.cu:44); callers that cannot
 * synchronise (graph replay) use this to report a truncated row on their next call. */
NNPOPS_API int nnpops_ani_overflow_poll(nnpops_ani_t h, int* flags, int* max_radial_neighbors, int* max_angular_neighbors);
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
/* One periodic box sharded over shard_count GPUs (one process per GPU): this model evaluates only the centres i with
 * i % shard_count == shard_rank, with ALL atoms as neighbour candidates (every rank passes the full position array).  energy is the
 * rank's partial sum and position_grad its partial dE/dx over all atoms; the caller sums both over the ranks (one all-reduce of
 * 4 + 12 * num_atoms bytes, e.g. ncclAllReduce over NVLink) -- the "degenerate halo" exchange of SURVEY.md section 8e.  The reference has no
 * multi-GPU path. */
NNPOPS_API int nnpops_ani_model_create_sharded(nnpops_ani_model_t* out, int num_atoms, int num_species, float radial_cutoff, float angular_cutoff,
                            const int* atom_species, int n_radial, const float* radial_fn, int n_angular, const float* angular_fn,
                            int ensemble_size, int num_layers, const int* dims, const float* params, int mlp_impl,
                            int max_radial_neighbors, int max_angular_neighbors, int shard_rank, int shard_count);
/* Spatial domain decomposition (one box over several GPUs, SURVEY.md section 8e variant i): the model of ONE rank's brick.  The atoms are
 * the brick's own atoms plus the ghost atoms within the radial cutoff around it; owned[i] != 0 marks the atoms whose AEV and network
 * this rank evaluates (host unsigned char [num_atoms]).  Energy and dE/dx are PARTIAL: the gradient rows of the ghosts go back to their
 * owners and the energies are summed (nnpops_b200.OptimizedTorchANI.HaloBoxANI does both over NCCL).  The reference is single-device. */
NNPOPS_API int nnpops_ani_model_create_owned(nnpops_ani_model_t* out, int num_atoms, int num_species, float radial_cutoff, float angular_cutoff,
                                  const int* atom_species, int n_radial, const float* radial_fn, int n_angular, const float* angular_fn,
                                  int ensemble_size, int num_layers, const int* dims, const float* params, int mlp_impl,
                                  int max_radial_neighbors, int max_angular_neighbors, const unsigned char* owned);
NNPOPS_API void nnpops_ani_model_destroy(nnpops_ani_model_t h);
/* energy: device float[1] (sum over atoms of the ensemble-mean atomic energies, no self-energy shift);
 * position_grad: device float [num_atoms][3] = dE/dx */
NNPOPS_API int nnpops_ani_model_energy_grad(nnpops_ani_model_t h, const float* positions, const float* box, float* energy,
                                 float* position_grad, void* stream);
/* same call with HOST buffers (pinned or pageable): copies in, evaluates, copies energy and gradient out, synchronises */
NNPOPS_API int nnpops_ani_model_energy_grad_host(nnpops_ani_model_t h, const float* positions_host, const float* box_host, float* energy_host,
                                      float* position_grad_host, void* stream);
/* introspection for tests and benchmarks: device pointers to the species-sorted AEV matrix (active columns only) and its gradient, the row
 * stride in floats, and (host, num_atoms ints) the row of each atom */
NNPOPS_API int nnpops_ani_model_buffers(nnpops_ani_model_t h, float** features, float** feature_grad, int* stride, int* row_of_atom);
/* copy the AEV matrix (which = 0) or its gradient (which = 1) of the last evaluation into out: device float [num_atoms][aev_length],
 * ATOM order, full AEV layout; columns of species absent from the system read 0 (their gradient is not formed) */
NNPOPS_API int nnpops_ani_model_read_features(nnpops_ani_model_t h, int which, float* out, void* stream);
NNPOPS_API int nnpops_ani_model_work(nnpops_ani_model_t h, long long* triples, long long* radial_pairs, double* mlp_flops_forward, void* stream);
/* mlp_flops_forward above counts the network on the model's full AEV length (the algorithmic figure).  The fused model evaluates
 * the AEV and the first layer only on the columns whose neighbour species occur in the system (the others are identically zero):
 * aev_length = full length, active_features = columns kept, mlp_flops_forward_executed = flops actually issued per forward. */
/* Verlet skin of the neighbour search (0 = off, the default): candidate rows within cutoff + skin are kept and reused until an atom has
 * moved more than skin / 2 or the box has changed; the decision is taken on the device in every call (no host synchronisation, works
 * inside CUDA graphs) and the rows the AEV kernels see are exact at every step.  The reference rebuilds its N x N neighbour table in
 * every call (CudaANISymmetryFunctions.cu:203-209); its callers' steady state is an MD loop (BenchmarkCudaCFConv.cu:105-112).
 * skin_stats synchronises and returns how many calls rebuilt / reused the candidate rows since set_skin. */
NNPOPS_API int nnpops_ani_set_skin(nnpops_ani_t h, float skin);
NNPOPS_API int nnpops_ani_model_set_skin(nnpops_ani_model_t h, float skin);
NNPOPS_API int nnpops_ani_model_skin_stats(nnpops_ani_model_t h, unsigned long long* rebuilds, unsigned long long* reuses);
/* *fused = 1 when the network runs as the one-kernel layer chain (csrc/mlp_chain.cu), 0 for the per-layer GEMM launches */
NNPOPS_API int nnpops_ani_model_mlp_fused(nnpops_ani_model_t h, int* fused);
NNPOPS_API int nnpops_ani_model_info(nnpops_ani_model_t h, int* aev_length, int* active_features, double* mlp_flops_forward_executed);
NNPOPS_API int nnpops_ani_model_overflowed(nnpops_ani_model_t h, int* flags);
NNPOPS_API int nnpops_ani_model_overflow_poll(nnpops_ani_model_t h, int* flags, int* max_radial_neighbors, int* max_angular_neighbors);
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

/* --------------------------------------------------------------------------------------------------------------------
 * getNeighborPairs.  Replaces torch op neighbors::getNeighborPairs (schema src/pytorch/neighbors/neighbors.cpp:4; kernels
 * getNeighborPairsCUDA.cu:31-195; semantics getNeighborPairsCPU.cpp:56-98 and getNeighborPairs.py:60-96).
 * positions: device [num_atoms][3]; box: device [3][3] or NULL; cutoff inclusive (distance <= cutoff);
 * max_num_pairs: -1 = all pairs (outputs of length num_atoms(num_atoms-1)/2, slot row(row-1)/2+col, misses -1 / NaN), or a
 * positive capacity (compacted list, padded with -1 / NaN, overflow dropped);
 * neighbors: device int32 [2][P] (row > col); deltas: device [P][3] = pos[row]-pos[col] (minimum image); distances: device [P];
 * num_found: device int32 [1] = number of pairs inside the cutoff (may exceed the capacity).  No host synchronisation:
 * CUDA-graph capturable after one warm-up call with the same num_atoms.
 * ------------------------------------------------------------------------------------------------------------------ */
NNPOPS_API int nnpops_neighbor_pairs_f32(const float* positions, const float* box, int num_atoms, float cutoff, long long max_num_pairs,
                                         int* neighbors, float* deltas, float* distances, int* num_found, void* stream);
NNPOPS_API int nnpops_neighbor_pairs_f64(const double* positions, const double* box, int num_atoms, double cutoff, long long max_num_pairs,
                                         int* neighbors, double* deltas, double* distances, int* num_found, void* stream);
/* backward (getNeighborPairsCUDA.cu:80-101,166-195): grad_positions [num_atoms][3] is overwritten */
NNPOPS_API int nnpops_neighbor_pairs_backward_f32(const int* neighbors, const float* deltas, const float* distances, const float* grad_deltas,
                                                  const float* grad_distances, long long num_pairs, int num_atoms, float* grad_positions,
                                                  void* stream);
NNPOPS_API int nnpops_neighbor_pairs_backward_f64(const int* neighbors, const double* deltas, const double* distances,
                                                  const double* grad_deltas, const double* grad_distances, long long num_pairs,
                                                  int num_atoms, double* grad_positions, void* stream);

/* --------------------------------------------------------------------------------------------------------------------
 * PME.  Replaces torch ops pme::pme_direct and pme::pme_reciprocal (schemas src/pytorch/pme/pme.cpp:3-6; CUDA
 * implementations pmeCUDA.cu:278-317 and :319-418).  All arrays device, fp32; exclusions int32 [num_atoms][max_exclusions] with
 * rows sorted descending and padded with -1 (as pme.py:92 prepares them).
 * pme_direct: energy [1]; pos_deriv [num_atoms][3] and charge_deriv [num_atoms] receive dE/dx and dE/dq (the reference
 *   computes them in forward and scales them in backward, pmeCUDA.cu:306,310-316).
 * pme_reciprocal_forward: energy [1]; recip_grid: float2 [gridx][gridy][gridz/2+1], the convolved half-complex grid to keep for
 *   backward (pmeCUDA.cu:365).  order 4 or 5 only, like the reference CUDA path (pmeCUDA.cu:347).
 * ------------------------------------------------------------------------------------------------------------------ */
NNPOPS_API int nnpops_pme_direct(const float* positions, const float* charges, const int* neighbors, const float* deltas,
                                 const float* distances, const int* exclusions, int num_atoms, long long num_pairs, int max_exclusions,
                                 float alpha, float coulomb, float* energy, float* pos_deriv, float* charge_deriv, void* stream);
/* Direct space WITHOUT a pair list: what PME.compute_direct does in two steps (pme.py:131-165: getNeighborPairs + pme::pme_direct;
 * kernels getNeighborPairsCUDA.cu:31-101 + pmeCUDA.cu:30-95) as one centre-owned traversal of a cell list -- same accepted pairs
 * (reference minimum-image arithmetic, distance <= cutoff), same exclusion test and correction, no 24-bytes-per-pair list and no
 * atomics.  box: [3][3].  shard_index / shard_count (0 / 1 for everything) select a contiguous range of the cell-sorted atoms, a
 * slab of the box, as the centres of this call: energy and derivatives of the shards add up to the whole (atoms outside the range
 * receive zeros), so one process per GPU evaluates one shard and the results are summed by an all-reduce. */
NNPOPS_API int nnpops_pme_direct_fused(const float* positions, const float* charges, const float* box, const int* exclusions,
                                       int num_atoms, int max_exclusions, float cutoff, float alpha, float coulomb, int shard_index,
                                       int shard_count, float* energy, float* pos_deriv, float* charge_deriv, void* stream);
NNPOPS_API int nnpops_pme_reciprocal_forward(const float* positions, const float* charges, const float* box, int num_atoms, int gridx,
                                             int gridy, int gridz, int order, float alpha, float coulomb, const float* xmoduli,
                                             const float* ymoduli, const float* zmoduli, float* energy, float* recip_grid, void* stream);
/* The forward pass in two stages, for a system sharded over several GPUs (one process per GPU): every rank spreads ITS atoms into a
 * full-size real grid float [gridx][gridy][gridz] (zeroed by the call), the ranks sum their grids (one all-reduce, e.g. ncclAllReduce
 * over NVLink), and every rank solves the same grid (FFT + convolution + energy) so that nnpops_pme_reciprocal_backward can then be
 * called with the rank's own atoms.  nnpops_pme_reciprocal_forward == spread into an internal grid + solve.  (pmeCUDA.cu:30-168.) */
NNPOPS_API int nnpops_pme_spread(const float* positions, const float* charges, const float* box, int num_atoms, int gridx, int gridy,
                                 int gridz, int order, float coulomb, float* real_grid, void* stream);
NNPOPS_API int nnpops_pme_solve(float* real_grid, const float* box, int gridx, int gridy, int gridz, float alpha, const float* xmoduli,
                                const float* ymoduli, const float* zmoduli, float* energy, float* recip_grid, void* stream);
NNPOPS_API int nnpops_pme_reciprocal_backward(const float* positions, const float* charges, const float* box, int num_atoms, int gridx,
                                              int gridy, int gridz, int order, float coulomb, const float* recip_grid, float* pos_deriv,
                                              float* charge_deriv, void* stream);

/* --------------------------------------------------------------------------------------------------------------------
 * SchNet CFConv.  Replaces classes CFConvNeighbors / CFConv (src/schnet/CFConv.h:37-85 and :109-217; CUDA implementations
 * src/schnet/CudaCFConv.cu:132-188, :352-378, :484-528) which the torch Holders drive from src/pytorch/CFConvNeighbors.cpp:39-75
 * and src/pytorch/CFConv.cpp:71-190.
 *   CFConvNeighbors(numAtoms, cutoff, periodic) + build(positions, box)         -> nnpops_cfconv_neighbors_create / _build
 *   CFConv(numAtoms, width, numGaussians, cutoff, periodic, gaussianWidth, activation, w1, b1, w2, b2)
 *                                                                                -> nnpops_cfconv_create
 *   compute(neighbors, positions, box, input, output)                           -> nnpops_cfconv_compute
 *   backprop(neighbors, positions, box, input, outputDeriv, inputDeriv, positionDeriv) -> nnpops_cfconv_backprop
 * w1: device float, indexed [width][num_gaussians] row-major exactly as the reference indexes the memory it is given
 * (CpuCFConv.cpp:160-168; the torch Holder passes a [num_gaussians, width] tensor's storage unchanged, CFConv.cpp:131-132);
 * w2: device [width][width]; b1, b2: device [width].  activation: 0 = shifted softplus, 1 = tanh (CFConv.h:111-120).
 * Periodicity is decided per build() by box != NULL.  positions/box of compute/backprop are those given to build() (the
 * list stores the sorted coordinates), so they are not passed again.  build() reads back one integer (the pair count) to size
 * the list.  points_per_sigma: resolution of the radial filter table (0 = default 32).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct nnpops_cfconv_neighbors* nnpops_cfconv_neighbors_t;
typedef struct nnpops_cfconv* nnpops_cfconv_t;
NNPOPS_API int nnpops_cfconv_neighbors_create(nnpops_cfconv_neighbors_t* out, int num_atoms, float cutoff);
NNPOPS_API void nnpops_cfconv_neighbors_destroy(nnpops_cfconv_neighbors_t h);
NNPOPS_API int nnpops_cfconv_neighbors_build(nnpops_cfconv_neighbors_t h, const float* positions, const float* box, void* stream);
NNPOPS_API int nnpops_cfconv_neighbors_num_pairs(nnpops_cfconv_neighbors_t h, long long* num_pairs);
NNPOPS_API int nnpops_cfconv_create(nnpops_cfconv_t* out, int width, int num_gaussians, float cutoff, float gaussian_width, int activation,
                                    const float* w1, const float* b1, const float* w2, const float* b2, int points_per_sigma);
NNPOPS_API void nnpops_cfconv_destroy(nnpops_cfconv_t h);
/* input, output: device float [num_atoms][width] */
NNPOPS_API int nnpops_cfconv_compute(nnpops_cfconv_t h, nnpops_cfconv_neighbors_t neighbors, const float* input, float* output, void* stream);
/* input_grad: device [num_atoms][width]; position_grad: device [num_atoms][3]; both overwritten */
NNPOPS_API int nnpops_cfconv_backprop(nnpops_cfconv_t h, nnpops_cfconv_neighbors_t neighbors, const float* input, const float* output_grad,
                                      float* input_grad, float* position_grad, void* stream);

/* development aid: mean milliseconds of one tcgen05 GEMM shape (C[m, batch*n] = A.B^T per batch member) on synthetic operands;
 * mode = epilogue (0 fp32, 1 bias+CELU, 2 celu' mask, 3 last hidden layer); streaming != 0 disables the resident-B variant */
/* measured fp32 FMA throughput of the current device in TFLOP/s (denominator of the AEV kernels' arithmetic roofline) */
NNPOPS_API int nnpops_debug_fma_peak(double* tflops);
NNPOPS_API int nnpops_debug_gemm_bench(int m, int n, int k, int batch, int mode, int streaming, int iters, double* ms);

#ifdef __cplusplus
}
#endif
#endif
