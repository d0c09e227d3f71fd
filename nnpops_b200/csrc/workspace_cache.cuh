// Scratch buffers of the stateless ops (getNeighborPairs, PME): one workspace per (device, shape, STREAM).
//
// The reference allocates its scratch per call through the torch allocator (getNeighborPairsCUDA.cu:120-127, pmeCUDA.cu:330-352); the
// C ABI has no allocator to borrow, so workspaces are cached.  Rules that keep the cache safe:
//   * the stream is part of the key, so calls that may run concurrently never share buffers (same stream = ordered);
//   * callers hold a shared_ptr for the duration of the call, so an entry cannot disappear under another thread;
//   * an entry that has been used while its stream was capturing a CUDA graph is pinned: replays write its addresses, it is never freed;
//   * eviction (least recently used, only un-pinned entries nobody holds) synchronises the entry's stream before its memory goes away.
#pragma once
#include <map>
#include <memory>
#include <mutex>
#include "common.cuh"

namespace nnpops {

struct WorkspaceBase {
    cudaStream_t stream = nullptr;
    bool pinned = false;                 // referenced by a captured graph
    unsigned long long lastUse = 0;
    virtual ~WorkspaceBase() {}
};

template <typename WS, typename Shape>
class WorkspaceCache {
public:
    explicit WorkspaceCache(size_t capacity) : cap_(capacity) {}
    // make: () -> WS* (allocates).  A miss while `stream` is capturing a CUDA graph first looks for a workspace of the same shape made
    // on another stream -- the torch idiom warms an op up on a side stream and captures on the graph's own -- and shares it (pinned);
    // only when there is none does it allocate, with the thread's capture mode relaxed for the duration as torch's allocator does.
    template <typename Make>
    std::shared_ptr<WS> get(const Shape& shape, cudaStream_t stream, Make make) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess) { cudaGetLastError(); cs = cudaStreamCaptureStatusNone; }
        const bool capturing = cs != cudaStreamCaptureStatusNone;
        std::lock_guard<std::mutex> lock(mu_);
        const Key key(shape, stream);
        auto it = map_.find(key);
        if (it == map_.end() && capturing) {
            for (auto jt = map_.begin(); jt != map_.end(); ++jt)
                if (jt->first.first == shape) { it = map_.emplace(key, jt->second).first; break; }
        }
        if (it == map_.end()) {
            if (map_.size() >= cap_) evict_one();
            cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
            if (capturing) cudaThreadExchangeStreamCaptureMode(&mode);
            std::shared_ptr<WS> ws;
            try {
                ws.reset(make());
            } catch (...) {
                if (capturing) cudaThreadExchangeStreamCaptureMode(&mode);
                throw;
            }
            if (capturing) cudaThreadExchangeStreamCaptureMode(&mode);
            ws->stream = stream;
            it = map_.emplace(key, ws).first;
        }
        it->second->lastUse = ++tick_;
        it->second->pinned = it->second->pinned || capturing;
        return it->second;
    }

private:
    typedef std::pair<Shape, cudaStream_t> Key;
    void evict_one() {
        auto victim = map_.end();
        for (auto it = map_.begin(); it != map_.end(); ++it)
            if (!it->second->pinned && it->second.use_count() == 1 && (victim == map_.end() || it->second->lastUse < victim->second->lastUse))
                victim = it;
        if (victim == map_.end()) return;   // everything is pinned or in use: grow instead
        if (cudaStreamSynchronize(victim->second->stream) != cudaSuccess) cudaGetLastError();   // e.g. the stream was destroyed: nothing in flight
        map_.erase(victim);
    }
    std::mutex mu_;
    std::map<Key, std::shared_ptr<WS>> map_;
    unsigned long long tick_ = 0;
    size_t cap_;
};

// multiprocessor count of the CURRENT device (a process may drive several)
inline int current_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    NNP_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) {
        int n = 0;
        NNP_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        return n;
    }
    if (!cached[dev]) NNP_CUDA_CHECK(cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev));
    return cached[dev];
}

}  // namespace nnpops
