// libNNPOpsPyTorch.so -- the torch custom-op surface of the reference, re-registered on top of the B200 C ABI
// (include/nnpops_b200.h).  Namespaces, class names, method names and schemas are those of the reference so that its Python
// wrappers (src/pytorch/*.py), TorchScript modules and openmm-torch call this library unchanged:
//   NNPOpsANISymmetryFunctions::{Holder, operation}     src/pytorch/SymmetryFunctions.cpp:265-284
//   NNPOpsBatchedNN::BatchedLinear                      src/pytorch/BatchedNN.cpp:44-50
//   NNPOpsCFConvNeighbors::Holder                       src/pytorch/CFConvNeighbors.cpp:77-85
//   NNPOpsCFConv::{Holder, operation}                   src/pytorch/CFConv.cpp:276-291
//   neighbors::getNeighborPairs                         src/pytorch/neighbors/neighbors.cpp:3-5
//   pme::{pme_direct, pme_reciprocal}                   src/pytorch/pme/pme.cpp:3-6
// Only CUDA tensors are accepted: there is no CPU implementation behind this library.
#include <torch/script.h>
#include <torch/serialize/archive.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <c10/cuda/CUDAGraphsC10Utils.h>
#include <sstream>
#include "../../../include/nnpops_b200.h"

namespace {

using torch::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::tensor_list;

void check(int rc) {
    if (rc != 0) throw std::runtime_error(nnpops_last_error());
}
void* stream_of(const Tensor& t) { return (void*)c10::cuda::getCurrentCUDAStream(t.get_device()).stream(); }
void require_cuda(const Tensor& t, const char* name) {
    TORCH_CHECK(t.is_cuda(), "nnpops_b200: \"", name, "\" has to be a CUDA tensor (there is no CPU fallback)");
}

// ---------------------------------------------------------------------------------------------- ANI symmetry functions
class AniHolder : public torch::CustomClassHolder {
public:
    AniHolder(int64_t numSpecies, double Rcr, double Rca, const std::vector<double>& EtaR, const std::vector<double>& ShfR,
              const std::vector<double>& EtaA, const std::vector<double>& Zeta, const std::vector<double>& ShfA,
              const std::vector<double>& ShfZ, const std::vector<int64_t>& atomSpecies)
        : numSpecies(numSpecies), Rcr(Rcr), Rca(Rca), EtaR(EtaR), ShfR(ShfR), EtaA(EtaA), Zeta(Zeta), ShfA(ShfA), ShfZ(ShfZ),
          atomSpecies(atomSpecies) {}
    ~AniHolder() override { nnpops_ani_destroy(impl); }

    tensor_list forward(const Tensor& positions, const c10::optional<Tensor>& cellOpt) {
        // same validation and messages as Holder::forward, SymmetryFunctions.cpp:76-102
        if (positions.scalar_type() != torch::kFloat32) throw std::runtime_error("The type of \"positions\" has to be float32");
        if (positions.dim() != 2) throw std::runtime_error("The shape of \"positions\" has to have 2 dimensions");
        if (positions.size(0) != (int64_t)atomSpecies.size())
            throw std::runtime_error("The size of the 1nd dimension of \"positions\" has to be " + std::to_string(atomSpecies.size()));
        if (positions.size(1) != 3) throw std::runtime_error("The size of the 2nd dimension of \"positions\" has to be 3");
        require_cuda(positions, "positions");
        Tensor cell;
        if (cellOpt) {
            cell = *cellOpt;
            if (cell.scalar_type() != torch::kFloat32) throw std::runtime_error("The type of \"cell\" has to be float32");
            if (cell.dim() != 2 || cell.size(0) != 3 || cell.size(1) != 3) throw std::runtime_error("The shape of \"cell\" has to be (3, 3)");
            if (cell.device() != positions.device()) throw std::runtime_error("\"cell\" has to be on the same device as \"positions\"");
            cell = cell.contiguous();
        }
        c10::cuda::CUDAGuard guard(positions.device());
        if (!impl) {
            device = positions.device();
            periodic = cellOpt.has_value();
            createImpl();
        }
        if (positions.device() != device) throw std::runtime_error("The device of \"positions\" has changed");
        const Tensor pos = positions.contiguous();
        const int64_t n = pos.size(0);
        const auto opt = torch::TensorOptions().device(device).dtype(torch::kFloat32);
        Tensor radial = torch::empty({n, numSpecies * nRadial}, opt);
        Tensor angular = torch::empty({n, numSpecies * (numSpecies + 1) / 2 * nAngular}, opt);
        // The reference keeps an N x N neighbour table and so has no neighbour limit (CudaANISymmetryFunctions.cu:44); here rows have a
        // capacity.  A row that overflowed is detected right after the call and the call repeated with larger rows, so the result is
        // never silently truncated.  While a CUDA graph is being captured nothing may synchronise: the check is skipped there and the
        // flag polled on the next eager call instead.
        const bool capturing = c10::cuda::currentStreamCaptureStatusMayInitCtx() != c10::cuda::CaptureStatus::None;
        if (!capturing) raiseIfOverflowedEarlier();
        for (int attempt = 0;; attempt++) {
            check(nnpops_ani_forward(impl, pos.data_ptr<float>(), cellOpt ? cell.data_ptr<float>() : nullptr, radial.data_ptr<float>(),
                                     angular.data_ptr<float>(), stream_of(pos)));
            if (capturing) break;
            int flags = 0;
            check(nnpops_ani_overflowed(impl, &flags));   // synchronises, like the reference's host read of the box (SymmetryFunctions.cpp:104-108)
            if (!flags) break;
            TORCH_CHECK(attempt < 8, "nnpops_b200: neighbour rows still overflow after growing them 8 times");
            if (flags & 1) maxRadial = 2 * (maxRadial > 0 ? maxRadial : 256);
            if (flags & 2) maxAngular = 2 * (maxAngular > 0 ? maxAngular : 64);
            nnpops_ani_destroy(impl);
            impl = nullptr;
            createImpl();
        }
        return {radial, angular};
    }

    // row capacities in use (0 = library default) and whether any row has overflowed so far (blocks)
    int64_t maxRadialNeighbors() const { return maxRadial; }
    int64_t maxAngularNeighbors() const { return maxAngular; }
    int64_t overflowed() {
        int flags = 0;
        if (impl) check(nnpops_ani_overflowed(impl, &flags));
        return flags;
    }

    tensor_list backward(const tensor_list& grads) {
        TORCH_CHECK(impl != nullptr, "backward() called before forward()");
        c10::cuda::CUDAGuard guard(device);
        const Tensor rg = grads[0].contiguous(), ag = grads[1].contiguous();
        Tensor positionsGrad = torch::empty({(int64_t)atomSpecies.size(), 3}, torch::TensorOptions().device(device).dtype(torch::kFloat32));
        check(nnpops_ani_backward(impl, rg.data_ptr<float>(), ag.data_ptr<float>(), positionsGrad.data_ptr<float>(), stream_of(rg)));
        return {Tensor(), positionsGrad, Tensor()};   // no gradient for the holder and the cell (SymmetryFunctions.cpp:174)
    }

    // state = the ten constructor arguments, same archive keys as the reference (SymmetryFunctions.cpp:177-218)
    std::string serialize() const {
        torch::serialize::OutputArchive ar;
        ar.write("numSpecies", numSpecies); ar.write("Rcr", Rcr); ar.write("Rca", Rca);
        ar.write("EtaR", EtaR); ar.write("ShfR", ShfR); ar.write("EtaA", EtaA); ar.write("Zeta", Zeta); ar.write("ShfA", ShfA);
        ar.write("ShfZ", ShfZ); ar.write("atomSpecies", atomSpecies);
        std::stringstream ss;
        ar.save_to(ss);
        return ss.str();
    }
    static c10::intrusive_ptr<AniHolder> deserialize(const std::string& state) {
        std::stringstream ss(state);
        torch::serialize::InputArchive ar;
        ar.load_from(ss, torch::kCPU);
        torch::IValue nS, rcr, rca, etaR, shfR, etaA, zeta, shfA, shfZ, sp;
        ar.read("numSpecies", nS); ar.read("Rcr", rcr); ar.read("Rca", rca); ar.read("EtaR", etaR); ar.read("ShfR", shfR);
        ar.read("EtaA", etaA); ar.read("Zeta", zeta); ar.read("ShfA", shfA); ar.read("ShfZ", shfZ); ar.read("atomSpecies", sp);
        return c10::make_intrusive<AniHolder>(nS.toInt(), rcr.toDouble(), rca.toDouble(), etaR.toDoubleVector(), shfR.toDoubleVector(),
                                              etaA.toDoubleVector(), zeta.toDoubleVector(), shfA.toDoubleVector(), shfZ.toDoubleVector(),
                                              sp.toIntVector());
    }

private:
    void createImpl() {
        std::vector<float> radialFn, angularFn;   // function order of SymmetryFunctions.cpp:110-120
        for (const float eta : EtaR)
            for (const float rs : ShfR) { radialFn.push_back(eta); radialFn.push_back(rs); }
        for (const float eta : EtaA)
            for (const float zeta : Zeta)
                for (const float rs : ShfA)
                    for (const float thetas : ShfZ) { angularFn.push_back(eta); angularFn.push_back(rs); angularFn.push_back(zeta); angularFn.push_back(thetas); }
        std::vector<int> species(atomSpecies.begin(), atomSpecies.end());
        nRadial = (int)radialFn.size() / 2; nAngular = (int)angularFn.size() / 4;
        check(nnpops_ani_create(&impl, (int)species.size(), (int)numSpecies, (float)Rcr, (float)Rca, species.data(), nRadial, radialFn.data(),
                                nAngular, angularFn.data(), 1, maxRadial, maxAngular));
    }
    // an overflow recorded by a call that could not check for itself (graph capture / replay)
    void raiseIfOverflowedEarlier() {
        int flags = 0;
        check(nnpops_ani_overflow_poll(impl, &flags, nullptr, nullptr));
        TORCH_CHECK(!flags, "nnpops_b200: a neighbour row overflowed during an earlier call that could not be checked (CUDA graph); "
                            "its AEVs were truncated. Run one eager forward before capturing so that the rows can grow.");
    }

    int64_t numSpecies;
    double Rcr, Rca;
    std::vector<double> EtaR, ShfR, EtaA, Zeta, ShfA, ShfZ;
    std::vector<int64_t> atomSpecies;
    torch::Device device = torch::kCPU;
    bool periodic = false;
    int nRadial = 0, nAngular = 0;
    int maxRadial = 0, maxAngular = 0;   // neighbour-row capacities handed to the library (0 = its defaults, 256 / 64); grown on overflow
    nnpops_ani_t impl = nullptr;
};

class AniFunction : public torch::autograd::Function<AniFunction> {
public:
    static tensor_list forward(AutogradContext* ctx, const c10::intrusive_ptr<AniHolder>& holder, const Tensor& positions,
                               const c10::optional<Tensor>& box) {
        ctx->saved_data["holder"] = holder;
        return holder->forward(positions, box);
    }
    static tensor_list backward(AutogradContext* ctx, const tensor_list& grads) {
        const auto holder = ctx->saved_data["holder"].toCustomClass<AniHolder>();
        ctx->saved_data.erase("holder");
        return holder->backward(grads);
    }
};

tensor_list ani_operation(const c10::optional<c10::intrusive_ptr<AniHolder>>& holder, const Tensor& positions,
                          const c10::optional<Tensor>& box) {
    return AniFunction::apply(*holder, positions, box);
}

// ---------------------------------------------------------------------------------------------- BatchedLinear
class BatchedLinearFunction : public torch::autograd::Function<BatchedLinearFunction> {
public:
    static Tensor forward(AutogradContext* ctx, const Tensor& vectors, const Tensor& weights, const Tensor& biases) {
        require_cuda(weights, "weights");
        TORCH_CHECK(weights.dim() == 5 && vectors.dim() == 5 && biases.dim() == 5, "BatchedLinear expects 5-D tensors");
        TORCH_CHECK(weights.size(0) == 1 && weights.scalar_type() == torch::kFloat32, "BatchedLinear: weights must be float32 [1, N, M, out, in]");
        // the reference is torch::matmul(weights, vectors) + biases (BatchedNN.cpp:34), which would broadcast over a leading batch of
        // molecules; this kernel evaluates one molecule, so refuse a batch instead of silently computing molecule 0 only
        TORCH_CHECK(vectors.size(0) == 1 && biases.size(0) == 1, "BatchedLinear: only one molecule per call is supported (leading dimension must be 1)");
        const int64_t N = weights.size(1), M = weights.size(2), nOut = weights.size(3), nIn = weights.size(4);
        TORCH_CHECK(vectors.size(1) == N && (vectors.size(2) == M || vectors.size(2) == 1) && vectors.size(3) == nIn && vectors.size(4) == 1,
                    "BatchedLinear: vectors must be [1, N, M or 1, in, 1]");
        c10::cuda::CUDAGuard guard(weights.device());
        const Tensor v = vectors.to(torch::kFloat32).contiguous(), w = weights.contiguous(), b = biases.expand({1, N, M, nOut, 1}).contiguous();
        Tensor out = torch::empty({1, N, M, nOut, 1}, w.options());
        check(nnpops_batched_linear_forward(v.data_ptr<float>(), w.data_ptr<float>(), b.data_ptr<float>(), out.data_ptr<float>(), (int)N, (int)M,
                                            (int)vectors.size(2), (int)nOut, (int)nIn, stream_of(w)));
        ctx->save_for_backward({w});
        ctx->saved_data["vecModels"] = vectors.size(2);
        return out;
    }
    static tensor_list backward(AutogradContext* ctx, const tensor_list& grads) {
        const Tensor w = ctx->get_saved_variables()[0];
        const int64_t N = w.size(1), M = w.size(2), nOut = w.size(3), nIn = w.size(4);
        c10::cuda::CUDAGuard guard(w.device());
        const Tensor go = grads[0].contiguous();
        Tensor g = torch::empty({1, N, M, nIn, 1}, w.options());
        check(nnpops_batched_linear_backward(go.data_ptr<float>(), w.data_ptr<float>(), g.data_ptr<float>(), (int)N, (int)M, (int)nOut, (int)nIn,
                                             stream_of(w)));
        if (ctx->saved_data["vecModels"].toInt() == 1 && M != 1) g = g.sum(2, /*keepdim=*/true);   // broadcast ensemble axis
        return {g, Tensor(), Tensor()};   // gradient w.r.t. the vectors only (BatchedNN.cpp:40)
    }
};
Tensor batched_linear(const Tensor& vectors, const Tensor& weights, const Tensor& biases) {
    return BatchedLinearFunction::apply(vectors, weights, biases);
}

// ---------------------------------------------------------------------------------------------- CFConv
class CFNeighborsHolder : public torch::CustomClassHolder {
public:
    explicit CFNeighborsHolder(double cutoff) : cutoff(cutoff) {}
    ~CFNeighborsHolder() override { nnpops_cfconv_neighbors_destroy(impl); }
    void build(const Tensor& positions) {
        if (positions.scalar_type() != torch::kFloat32) throw std::runtime_error("The type of \"positions\" has to be float32");
        if (positions.dim() != 2 || positions.size(1) != 3) throw std::runtime_error("The shape of \"positions\" has to be (numAtoms, 3)");
        require_cuda(positions, "positions");
        c10::cuda::CUDAGuard guard(positions.device());
        if (!impl) {
            numAtoms = positions.size(0);
            device = positions.device();
            check(nnpops_cfconv_neighbors_create(&impl, (int)numAtoms, (float)cutoff));
        }
        if (positions.size(0) != numAtoms) throw std::runtime_error("The size of the 1nd dimension of \"positions\" has to be " + std::to_string(numAtoms));
        if (positions.device() != device) throw std::runtime_error("The device of \"positions\" has changed");
        const Tensor pos = positions.contiguous();
        // the reference torch surface is non-periodic (CFConvNeighbors.cpp:52,57,74)
        check(nnpops_cfconv_neighbors_build(impl, pos.data_ptr<float>(), nullptr, stream_of(pos)));
    }
    double cutoff;
    int64_t numAtoms = -1;
    torch::Device device = torch::kCPU;
    nnpops_cfconv_neighbors_t impl = nullptr;
};

class CFConvHolder : public torch::CustomClassHolder {
public:
    CFConvHolder(double gaussianWidth, const std::string& activation, const Tensor& weights1, const Tensor& biases1, const Tensor& weights2,
                 const Tensor& biases2)
        : gaussianWidth(gaussianWidth), activation(activation), weights1(weights1), biases1(biases1), weights2(weights2), biases2(biases2) {
        if (activation != "ssp" && activation != "tanh") throw std::invalid_argument("Invalid value of \"activation\"");
    }
    ~CFConvHolder() override { nnpops_cfconv_destroy(impl); }

    Tensor forward(const c10::IValue& neighborsValue, const Tensor& positions, const Tensor& input) {
        neighbors = neighborsValue.toCustomClass<CFNeighborsHolder>();
        TORCH_CHECK(neighbors->impl != nullptr, "CFConvNeighbors.build() has to be called before CFConv");
        if (positions.scalar_type() != torch::kFloat32) throw std::runtime_error("The type of \"positions\" has to be float32");
        if (input.scalar_type() != torch::kFloat32) throw std::runtime_error("The type of \"input\" has to be float32");
        require_cuda(input, "input");
        const int64_t W = weights1.size(1), G = weights1.size(0);
        // shape checks of CFConv.cpp:107-128
        TORCH_CHECK(weights1.dim() == 2 && biases1.dim() == 1 && biases1.size(0) == W, "inconsistent shapes of weights1/biases1");
        TORCH_CHECK(weights2.dim() == 2 && weights2.size(0) == W && weights2.size(1) == W && biases2.dim() == 1 && biases2.size(0) == W,
                    "inconsistent shapes of weights2/biases2");
        TORCH_CHECK(input.dim() == 2 && input.size(0) == neighbors->numAtoms && input.size(1) == W, "The shape of \"input\" has to be (numAtoms, numFilters)");
        c10::cuda::CUDAGuard guard(input.device());
        if (!impl) {
            const auto opt = torch::TensorOptions().device(input.device()).dtype(torch::kFloat32);
            // storage handed over unchanged and indexed [numFilters][numGaussians] by the kernels (CFConv.cpp:131-132)
            const Tensor w1 = weights1.to(opt).contiguous(), b1 = biases1.to(opt).contiguous(), w2 = weights2.to(opt).contiguous(),
                         b2 = biases2.to(opt).contiguous();
            check(nnpops_cfconv_create(&impl, (int)W, (int)G, (float)neighbors->cutoff, (float)gaussianWidth, activation == "ssp" ? 0 : 1,
                                       w1.data_ptr<float>(), b1.data_ptr<float>(), w2.data_ptr<float>(), b2.data_ptr<float>(), 0));
        }
        savedInput = input.contiguous();
        Tensor output = torch::empty_like(savedInput);
        check(nnpops_cfconv_compute(impl, neighbors->impl, savedInput.data_ptr<float>(), output.data_ptr<float>(), stream_of(savedInput)));
        return output;
    }
    tensor_list backward(const tensor_list& grads) {
        c10::cuda::CUDAGuard guard(savedInput.device());
        const Tensor go = grads[0].contiguous();
        Tensor inputGrad = torch::empty_like(savedInput);
        Tensor positionsGrad = torch::empty({savedInput.size(0), 3}, savedInput.options());
        check(nnpops_cfconv_backprop(impl, neighbors->impl, savedInput.data_ptr<float>(), go.data_ptr<float>(), inputGrad.data_ptr<float>(),
                                     positionsGrad.data_ptr<float>(), stream_of(go)));
        return {Tensor(), Tensor(), positionsGrad, inputGrad};   // CFConv.cpp:189
    }
    double gaussianWidth;
    std::string activation;
    Tensor weights1, biases1, weights2, biases2;

private:
    c10::intrusive_ptr<CFNeighborsHolder> neighbors;
    Tensor savedInput;
    nnpops_cfconv_t impl = nullptr;
};

class CFConvFunction : public torch::autograd::Function<CFConvFunction> {
public:
    static Tensor forward(AutogradContext* ctx, const c10::intrusive_ptr<CFConvHolder>& holder, const c10::IValue& neighbors,
                          const Tensor& positions, const Tensor& input) {
        ctx->saved_data["holder"] = holder;
        return holder->forward(neighbors, positions, input);
    }
    static tensor_list backward(AutogradContext* ctx, const tensor_list& grads) {
        const auto holder = ctx->saved_data["holder"].toCustomClass<CFConvHolder>();
        ctx->saved_data.erase("holder");
        return holder->backward(grads);
    }
};
Tensor cfconv_operation(const c10::optional<c10::intrusive_ptr<CFConvHolder>>& holder, const c10::IValue& neighbors, const Tensor& positions,
                        const Tensor& input) {
    return CFConvFunction::apply(*holder, neighbors, positions, input);
}

// ---------------------------------------------------------------------------------------------- getNeighborPairs
class NeighborFunction : public torch::autograd::Function<NeighborFunction> {
public:
    static tensor_list forward(AutogradContext* ctx, const Tensor& positions, const torch::Scalar& cutoff, const torch::Scalar& maxNumPairs,
                               const Tensor& boxVectors, bool checkErrors) {
        TORCH_CHECK(positions.dim() == 2, "Expected \"positions\" to have two dimensions");
        TORCH_CHECK(positions.size(0) > 0, "Expected the 1nd dimension size of \"positions\" to be more than 0");
        TORCH_CHECK(positions.size(1) == 3, "Expected the 2nd dimension size of \"positions\" to be 3");
        TORCH_CHECK(positions.is_contiguous(), "Expected \"positions\" to be contiguous");
        TORCH_CHECK(cutoff.to<double>() > 0, "Expected \"cutoff\" to be positive");
        const int64_t maxPairs = maxNumPairs.to<int64_t>();
        TORCH_CHECK(maxPairs > 0 || maxPairs == -1, "Expected \"max_num_pairs\" to be positive or equal to -1");
        const bool periodic = boxVectors.size(0) != 0;
        if (periodic) TORCH_CHECK(boxVectors.dim() == 2 && boxVectors.size(0) == 3 && boxVectors.size(1) == 3, "Expected \"box_vectors\" to have shape (3, 3)");
        c10::cuda::CUDAGuard guard(positions.device());
        const int64_t n = positions.size(0);
        const int64_t numPairs = maxPairs == -1 ? n * (n - 1) / 2 : maxPairs;
        const auto opt = positions.options();
        Tensor neighbors = torch::empty({2, numPairs}, opt.dtype(torch::kInt32));
        Tensor deltas = torch::empty({numPairs, 3}, opt), distances = torch::empty({numPairs}, opt);
        Tensor found = torch::empty({1}, opt.dtype(torch::kInt32));
        const Tensor box = periodic ? boxVectors.to(opt).contiguous() : Tensor();
        if (positions.scalar_type() == torch::kFloat32) {
            check(nnpops_neighbor_pairs_f32(positions.data_ptr<float>(), periodic ? box.data_ptr<float>() : nullptr, (int)n, cutoff.to<float>(),
                                            maxPairs, neighbors.data_ptr<int>(), deltas.data_ptr<float>(), distances.data_ptr<float>(),
                                            found.data_ptr<int>(), stream_of(positions)));
        } else if (positions.scalar_type() == torch::kFloat64) {
            check(nnpops_neighbor_pairs_f64(positions.data_ptr<double>(), periodic ? box.data_ptr<double>() : nullptr, (int)n, cutoff.to<double>(),
                                            maxPairs, neighbors.data_ptr<int>(), deltas.data_ptr<double>(), distances.data_ptr<double>(),
                                            found.data_ptr<int>(), stream_of(positions)));
        } else {
            TORCH_CHECK(false, "Expected \"positions\" to be float32 or float64");
        }
        if (checkErrors && maxPairs != -1)   // synchronises, like getNeighborPairsCUDA.cu:157-160
            TORCH_CHECK(found.item<int>() <= maxPairs, "The maximum number of pairs has been exceed! Increase \"max_num_pairs\"");
        ctx->save_for_backward({neighbors, deltas, distances});
        ctx->saved_data["num_atoms"] = n;
        ctx->mark_non_differentiable({neighbors, found});
        return {neighbors, deltas, distances, found};
    }
    static tensor_list backward(AutogradContext* ctx, const tensor_list& grads) {
        const auto saved = ctx->get_saved_variables();
        const Tensor neighbors = saved[0], deltas = saved[1], distances = saved[2];
        const int64_t n = ctx->saved_data["num_atoms"].toInt();
        c10::cuda::CUDAGuard guard(deltas.device());
        const Tensor gd = grads[1].defined() ? grads[1].contiguous() : torch::zeros_like(deltas);
        const Tensor gr = grads[2].defined() ? grads[2].contiguous() : torch::zeros_like(distances);
        Tensor gp = torch::empty({n, 3}, deltas.options());
        if (deltas.scalar_type() == torch::kFloat32)
            check(nnpops_neighbor_pairs_backward_f32(neighbors.data_ptr<int>(), deltas.data_ptr<float>(), distances.data_ptr<float>(),
                                                     gd.data_ptr<float>(), gr.data_ptr<float>(), distances.size(0), (int)n, gp.data_ptr<float>(),
                                                     stream_of(deltas)));
        else
            check(nnpops_neighbor_pairs_backward_f64(neighbors.data_ptr<int>(), deltas.data_ptr<double>(), distances.data_ptr<double>(),
                                                     gd.data_ptr<double>(), gr.data_ptr<double>(), distances.size(0), (int)n, gp.data_ptr<double>(),
                                                     stream_of(deltas)));
        return {gp, Tensor(), Tensor(), Tensor(), Tensor()};
    }
};

// ---------------------------------------------------------------------------------------------- PME
class PmeDirectFunction : public torch::autograd::Function<PmeDirectFunction> {
public:
    static Tensor forward(AutogradContext* ctx, const Tensor& positions, const Tensor& charges, const Tensor& neighbors, const Tensor& deltas,
                          const Tensor& distances, const Tensor& exclusions, const torch::Scalar& alpha, const torch::Scalar& coulomb) {
        c10::cuda::CUDAGuard guard(positions.device());
        const auto f32 = positions.options().dtype(torch::kFloat32);
        const Tensor pos = positions.to(f32).contiguous(), q = charges.to(f32).contiguous(), d = deltas.to(f32).contiguous(),
                     r = distances.to(f32).contiguous();
        const Tensor nb = neighbors.to(positions.options().dtype(torch::kInt32)).contiguous();
        const Tensor ex = exclusions.to(positions.options().dtype(torch::kInt32)).contiguous();
        const int64_t n = q.size(0);
        Tensor energy = torch::empty({}, f32), posDeriv = torch::empty({n, 3}, f32), chargeDeriv = torch::empty({n}, f32);
        const int maxExcl = ex.dim() == 2 ? (int)ex.size(1) : 0;
        check(nnpops_pme_direct(pos.data_ptr<float>(), q.data_ptr<float>(), nb.data_ptr<int>(), d.data_ptr<float>(), r.data_ptr<float>(),
                                maxExcl > 0 ? ex.data_ptr<int>() : nullptr, (int)n, nb.size(1), maxExcl, alpha.to<float>(), coulomb.to<float>(),
                                energy.data_ptr<float>(), posDeriv.data_ptr<float>(), chargeDeriv.data_ptr<float>(), stream_of(pos)));
        ctx->save_for_backward({posDeriv, chargeDeriv});
        return energy;
    }
    static tensor_list backward(AutogradContext* ctx, const tensor_list& grads) {
        const auto saved = ctx->get_saved_variables();
        Tensor none;
        return {saved[0] * grads[0], saved[1] * grads[0], none, none, none, none, none, none};
    }
};

class PmeReciprocalFunction : public torch::autograd::Function<PmeReciprocalFunction> {
public:
    static Tensor forward(AutogradContext* ctx, const Tensor& positions, const Tensor& charges, const Tensor& boxVectors, const torch::Scalar& gridx,
                          const torch::Scalar& gridy, const torch::Scalar& gridz, const torch::Scalar& order, const torch::Scalar& alpha,
                          const torch::Scalar& coulomb, const Tensor& xmoduli, const Tensor& ymoduli, const Tensor& zmoduli) {
        c10::cuda::CUDAGuard guard(positions.device());
        const auto f32 = positions.options().dtype(torch::kFloat32);
        const Tensor pos = positions.to(f32).contiguous(), q = charges.to(f32).contiguous(), box = boxVectors.to(f32).contiguous();
        const Tensor xm = xmoduli.to(f32).contiguous(), ym = ymoduli.to(f32).contiguous(), zm = zmoduli.to(f32).contiguous();
        const int gx = gridx.to<int>(), gy = gridy.to<int>(), gz = gridz.to<int>();
        Tensor energy = torch::empty({}, f32), recip = torch::empty({gx, gy, gz / 2 + 1, 2}, f32);
        check(nnpops_pme_reciprocal_forward(pos.data_ptr<float>(), q.data_ptr<float>(), box.data_ptr<float>(), (int)q.size(0), gx, gy, gz,
                                            order.to<int>(), alpha.to<float>(), coulomb.to<float>(), xm.data_ptr<float>(), ym.data_ptr<float>(),
                                            zm.data_ptr<float>(), energy.data_ptr<float>(), recip.data_ptr<float>(), stream_of(pos)));
        ctx->save_for_backward({pos, q, box, recip});
        ctx->saved_data["gridx"] = (int64_t)gx; ctx->saved_data["gridy"] = (int64_t)gy; ctx->saved_data["gridz"] = (int64_t)gz;
        ctx->saved_data["order"] = (int64_t)order.to<int>(); ctx->saved_data["coulomb"] = coulomb.to<double>();
        return energy;
    }
    static tensor_list backward(AutogradContext* ctx, const tensor_list& grads) {
        const auto saved = ctx->get_saved_variables();
        const Tensor pos = saved[0], q = saved[1], box = saved[2], recip = saved[3];
        c10::cuda::CUDAGuard guard(pos.device());
        Tensor posDeriv = torch::empty_like(pos), chargeDeriv = torch::empty_like(q);
        check(nnpops_pme_reciprocal_backward(pos.data_ptr<float>(), q.data_ptr<float>(), box.data_ptr<float>(), (int)q.size(0),
                                             (int)ctx->saved_data["gridx"].toInt(), (int)ctx->saved_data["gridy"].toInt(),
                                             (int)ctx->saved_data["gridz"].toInt(), (int)ctx->saved_data["order"].toInt(),
                                             (float)ctx->saved_data["coulomb"].toDouble(), recip.data_ptr<float>(), posDeriv.data_ptr<float>(),
                                             chargeDeriv.data_ptr<float>(), stream_of(pos)));
        Tensor none;
        return {posDeriv * grads[0], chargeDeriv * grads[0], none, none, none, none, none, none, none, none, none, none};
    }
};

// ---------------------------------------------------------------------------------------------- fused ANI model (extension)
// Not part of the reference surface: the scalable sibling of OptimizedTorchANI (SURVEY.md section 8b "mismatch list" / 8f.2) as a
// TorchScript-able custom class, so that a scripted module (e.g. for openmm-torch) can use the species-grouped tensor-core path.
class FusedAniHolder : public torch::CustomClassHolder {
public:
    FusedAniHolder(int64_t numSpecies, double Rcr, double Rca, const std::vector<double>& EtaR, const std::vector<double>& ShfR,
                   const std::vector<double>& EtaA, const std::vector<double>& Zeta, const std::vector<double>& ShfA,
                   const std::vector<double>& ShfZ, const std::vector<int64_t>& atomSpecies, int64_t ensembleSize,
                   const std::vector<int64_t>& dims, const Tensor& params, int64_t mlpImpl)
        : numSpecies(numSpecies), Rcr(Rcr), Rca(Rca), EtaR(EtaR), ShfR(ShfR), EtaA(EtaA), Zeta(Zeta), ShfA(ShfA), ShfZ(ShfZ),
          atomSpecies(atomSpecies), ensembleSize(ensembleSize), dims(dims), params(params.detach().to(torch::kCPU, torch::kFloat32).contiguous()),
          mlpImpl(mlpImpl) {
        TORCH_CHECK(numSpecies > 0 && dims.size() % (size_t)numSpecies == 0 && dims.size() / numSpecies >= 3,
                    "dims must hold numSpecies rows of numLayers+1 layer sizes");
    }
    ~FusedAniHolder() override { nnpops_ani_model_destroy(impl); }

    // -> {energy [1], dE/dpositions [N, 3]}
    tensor_list evaluate(const Tensor& positions, const c10::optional<Tensor>& cellOpt) {
        if (positions.scalar_type() != torch::kFloat32) throw std::runtime_error("The type of \"positions\" has to be float32");
        if (positions.dim() != 2 || positions.size(0) != (int64_t)atomSpecies.size() || positions.size(1) != 3)
            throw std::runtime_error("The shape of \"positions\" has to be (" + std::to_string(atomSpecies.size()) + ", 3)");
        require_cuda(positions, "positions");
        c10::cuda::CUDAGuard guard(positions.device());
        Tensor cell;
        if (cellOpt) {
            cell = cellOpt->to(positions.options()).contiguous();
            TORCH_CHECK(cell.dim() == 2 && cell.size(0) == 3 && cell.size(1) == 3, "The shape of \"cell\" has to be (3, 3)");
        }
        if (!impl) {
            device = positions.device();
            std::vector<float> radialFn, angularFn;
            for (const float eta : EtaR)
                for (const float rs : ShfR) { radialFn.push_back(eta); radialFn.push_back(rs); }
            for (const float eta : EtaA)
                for (const float zeta : Zeta)
                    for (const float rs : ShfA)
                        for (const float thetas : ShfZ) { angularFn.push_back(eta); angularFn.push_back(rs); angularFn.push_back(zeta); angularFn.push_back(thetas); }
            std::vector<int> species(atomSpecies.begin(), atomSpecies.end()), d(dims.begin(), dims.end());
            const int numLayers = (int)(dims.size() / numSpecies) - 1;
            check(nnpops_ani_model_create(&impl, (int)species.size(), (int)numSpecies, (float)Rcr, (float)Rca, species.data(), (int)radialFn.size() / 2,
                                          radialFn.data(), (int)angularFn.size() / 4, angularFn.data(), (int)ensembleSize, numLayers, d.data(),
                                          params.data_ptr<float>(), (int)mlpImpl, 0, 0));
        }
        if (positions.device() != device) throw std::runtime_error("The device of \"positions\" has changed");
        {   // never blocks: a row that overflowed in an earlier evaluation is reported now (the fused path must stay asynchronous)
            int flags = 0, capR = 0, capA = 0;
            check(nnpops_ani_model_overflow_poll(impl, &flags, &capR, &capA));
            TORCH_CHECK(!flags, "nnpops_b200: a neighbour row overflowed in an earlier evaluation (capacities ", capR, " radial / ", capA,
                        " angular): energies and forces since then are wrong. The system is denser than the rows allow.");
        }
        const Tensor pos = positions.contiguous();
        Tensor energy = torch::empty({1}, pos.options()), grad = torch::empty_like(pos);
        check(nnpops_ani_model_energy_grad(impl, pos.data_ptr<float>(), cellOpt ? cell.data_ptr<float>() : nullptr, energy.data_ptr<float>(),
                                           grad.data_ptr<float>(), stream_of(pos)));
        return {energy, grad};
    }

    using State = std::tuple<int64_t, double, double, std::vector<double>, std::vector<double>, std::vector<double>, std::vector<double>,
                             std::vector<double>, std::vector<double>, std::vector<int64_t>, int64_t, std::vector<int64_t>, Tensor, int64_t>;
    State state() const {
        return State(numSpecies, Rcr, Rca, EtaR, ShfR, EtaA, Zeta, ShfA, ShfZ, atomSpecies, ensembleSize, dims, params, mlpImpl);
    }

private:
    int64_t numSpecies;
    double Rcr, Rca;
    std::vector<double> EtaR, ShfR, EtaA, Zeta, ShfA, ShfZ;
    std::vector<int64_t> atomSpecies;
    int64_t ensembleSize;
    std::vector<int64_t> dims;
    Tensor params;
    int64_t mlpImpl;
    torch::Device device = torch::kCPU;
    nnpops_ani_model_t impl = nullptr;
};

class FusedAniFunction : public torch::autograd::Function<FusedAniFunction> {
public:
    static Tensor forward(AutogradContext* ctx, const c10::intrusive_ptr<FusedAniHolder>& holder, const Tensor& positions,
                          const c10::optional<Tensor>& cell) {
        const tensor_list r = holder->evaluate(positions, cell);
        ctx->save_for_backward({r[1]});
        return r[0];
    }
    static tensor_list backward(AutogradContext* ctx, const tensor_list& grads) {
        const Tensor g = ctx->get_saved_variables()[0];
        return {Tensor(), g * grads[0], Tensor()};
    }
};
Tensor fused_ani_operation(const c10::optional<c10::intrusive_ptr<FusedAniHolder>>& holder, const Tensor& positions,
                           const c10::optional<Tensor>& cell) {
    return FusedAniFunction::apply(*holder, positions, cell);
}

}  // namespace

TORCH_LIBRARY(NNPOpsFusedANI, m) {
    m.class_<FusedAniHolder>("Holder")
        .def(torch::init<int64_t, double, double, const std::vector<double>&, const std::vector<double>&, const std::vector<double>&,
                         const std::vector<double>&, const std::vector<double>&, const std::vector<double>&, const std::vector<int64_t>&, int64_t,
                         const std::vector<int64_t>&, const Tensor&, int64_t>())
        .def("evaluate", &FusedAniHolder::evaluate)
        .def_pickle([](const c10::intrusive_ptr<FusedAniHolder>& self) -> FusedAniHolder::State { return self->state(); },
                    [](FusedAniHolder::State st) -> c10::intrusive_ptr<FusedAniHolder> {
                        return c10::make_intrusive<FusedAniHolder>(std::get<0>(st), std::get<1>(st), std::get<2>(st), std::get<3>(st), std::get<4>(st),
                                                                   std::get<5>(st), std::get<6>(st), std::get<7>(st), std::get<8>(st), std::get<9>(st),
                                                                   std::get<10>(st), std::get<11>(st), std::get<12>(st), std::get<13>(st));
                    });
    m.def("operation", fused_ani_operation);
}

TORCH_LIBRARY(NNPOpsANISymmetryFunctions, m) {
    m.class_<AniHolder>("Holder")
        .def(torch::init<int64_t, double, double, const std::vector<double>&, const std::vector<double>&, const std::vector<double>&,
                         const std::vector<double>&, const std::vector<double>&, const std::vector<double>&, const std::vector<int64_t>&>())
        .def("forward", &AniHolder::forward)
        .def("backward", &AniHolder::backward)
        .def("max_radial_neighbors", &AniHolder::maxRadialNeighbors)
        .def("max_angular_neighbors", &AniHolder::maxAngularNeighbors)
        .def("overflowed", &AniHolder::overflowed)
        .def_pickle([](const c10::intrusive_ptr<AniHolder>& self) -> std::string { return self->serialize(); },
                    [](const std::string& state) -> c10::intrusive_ptr<AniHolder> { return AniHolder::deserialize(state); });
    m.def("operation", ani_operation);
}

TORCH_LIBRARY(NNPOpsBatchedNN, m) { m.def("BatchedLinear", batched_linear); }

TORCH_LIBRARY(NNPOpsCFConvNeighbors, m) {
    m.class_<CFNeighborsHolder>("Holder")
        .def(torch::init<double>())
        .def("build", &CFNeighborsHolder::build)
        .def_pickle([](const c10::intrusive_ptr<CFNeighborsHolder>& self) -> double { return self->cutoff; },
                    [](double cutoff) -> c10::intrusive_ptr<CFNeighborsHolder> { return c10::make_intrusive<CFNeighborsHolder>(cutoff); });
}

TORCH_LIBRARY(NNPOpsCFConv, m) {
    m.class_<CFConvHolder>("Holder")
        .def(torch::init<double, const std::string&, const Tensor&, const Tensor&, const Tensor&, const Tensor&>())
        .def("forward", &CFConvHolder::forward)
        .def("backward", &CFConvHolder::backward)
        // the state is the string the reference writes (CFConv.cpp:191-225: an OutputArchive with these six keys), so archives saved by
        // either library load on the other
        .def_pickle(
            [](const c10::intrusive_ptr<CFConvHolder>& self) -> std::string {
                torch::serialize::OutputArchive ar;
                ar.write("gaussianWidth", self->gaussianWidth); ar.write("activation", self->activation);
                ar.write("weights1", self->weights1.cpu()); ar.write("biases1", self->biases1.cpu());
                ar.write("weights2", self->weights2.cpu()); ar.write("biases2", self->biases2.cpu());
                std::stringstream ss;
                ar.save_to(ss);
                return ss.str();
            },
            [](const std::string& state) -> c10::intrusive_ptr<CFConvHolder> {
                std::stringstream ss(state);
                torch::serialize::InputArchive ar;
                ar.load_from(ss, torch::kCPU);
                torch::IValue gaussianWidth, activation;
                Tensor w1, b1, w2, b2;
                ar.read("gaussianWidth", gaussianWidth); ar.read("activation", activation);
                ar.read("weights1", w1); ar.read("biases1", b1); ar.read("weights2", w2); ar.read("biases2", b2);
                return c10::make_intrusive<CFConvHolder>(gaussianWidth.toDouble(), activation.toStringRef(), w1, b1, w2, b2);
            });
    m.def("operation", cfconv_operation);
}

TORCH_LIBRARY(neighbors, m) {
    m.def("getNeighborPairs(Tensor positions, Scalar cutoff, Scalar max_num_neighbors, Tensor box_vectors, bool checkErrors) -> "
          "(Tensor neighbors, Tensor deltas, Tensor distances, Tensor num_pairs)");
}
// The same entry serves the AutogradCUDA and the CUDA key: torch::autograd::Function::apply records a graph only when one is needed
// and the forward calls the C ABI directly (no re-dispatch), so calls under torch.inference_mode() / below autograd find a kernel too,
// as they do with the reference's CUDA registration (getNeighborPairsCUDA.cu:198-206).
#define NNPOPS_IMPL_BOTH(NS, BODY)                 \
    TORCH_LIBRARY_IMPL(NS, AutogradCUDA, m) { BODY } \
    TORCH_LIBRARY_IMPL(NS, CUDA, m) { BODY }

NNPOPS_IMPL_BOTH(neighbors,
    m.impl("getNeighborPairs", [](const Tensor& positions, const torch::Scalar& cutoff, const torch::Scalar& maxNumPairs, const Tensor& boxVectors,
                                  bool checkErrors) {
        const tensor_list r = NeighborFunction::apply(positions, cutoff, maxNumPairs, boxVectors, checkErrors);
        return std::make_tuple(r[0], r[1], r[2], r[3]);
    });
)

TORCH_LIBRARY(pme, m) {
    m.def("pme_direct(Tensor positions, Tensor charges, Tensor neighbors, Tensor deltas, Tensor distances, Tensor exclusions, Scalar alpha, "
          "Scalar coulomb) -> Tensor");
    m.def("pme_reciprocal(Tensor positions, Tensor charges, Tensor box_vectors, Scalar gridx, Scalar gridy, Scalar gridz, Scalar order, "
          "Scalar alpha, Scalar coulomb, Tensor xmoduli, Tensor ymoduli, Tensor zmoduli) -> Tensor");
}
NNPOPS_IMPL_BOTH(pme,
    m.impl("pme_direct", [](const Tensor& positions, const Tensor& charges, const Tensor& neighbors, const Tensor& deltas, const Tensor& distances,
                            const Tensor& exclusions, const torch::Scalar& alpha, const torch::Scalar& coulomb) {
        return PmeDirectFunction::apply(positions, charges, neighbors, deltas, distances, exclusions, alpha, coulomb);
    });
    m.impl("pme_reciprocal", [](const Tensor& positions, const Tensor& charges, const Tensor& boxVectors, const torch::Scalar& gridx,
                                const torch::Scalar& gridy, const torch::Scalar& gridz, const torch::Scalar& order, const torch::Scalar& alpha,
                                const torch::Scalar& coulomb, const Tensor& xmoduli, const Tensor& ymoduli, const Tensor& zmoduli) {
        return PmeReciprocalFunction::apply(positions, charges, boxVectors, gridx, gridy, gridz, order, alpha, coulomb, xmoduli, ymoduli, zmoduli);
    });
)
