// kdbuild_kernels.inl — the passes of the device KD-tree build as kernels (algorithm and per-item bodies: kdbuild.h; driver:
// csrc/kd_device_build.cpp). Included by launch_cuda.cu inside namespace hxr::dev.
//
// Per level of the tree: k_kd_bin (histograms of the big nodes: per reference, privatised in shared memory for the node a
// block starts in), k_kd_choose (ONE WARP PER NODE: lane k evaluates binned plane k, or the lanes share the bound edges of a
// small node held in shared memory; warp-shuffle argmin), k_kd_classify, two scans (CUB), k_kd_plan, three scans over the
// nodes, k_kd_emit, k_kd_scatter. HBM-bound streaming passes over the level's references except for the gathers of the
// triangle bounds (SoA doubles) through the reference's triangle index.

__global__ void __launch_bounds__(256) k_kd_bounds(const double* __restrict__ vertices, const int32_t* __restrict__ triV, uint32_t nTris, double* tb)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nTris; t += stride) kdb::bounds_item(t, vertices, triV, tb, nTris);
}

__global__ void __launch_bounds__(256) k_kd_iota(uint32_t* refTri, uint32_t* refNode, uint32_t n)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) { refTri[i] = i; refNode[i] = 0; }
}

#define HXR_KDB_HIST (3 * 2 * HXR_KDB_BINS)
#define HXR_KDB_CHUNK 2048u

__global__ void __launch_bounds__(256) k_kd_bin(kdb::Params P, const kdb::NodeWork* __restrict__ work, const uint32_t* __restrict__ refTri,
                                                const uint32_t* __restrict__ refNode, uint32_t nRefs, const double* __restrict__ tb, uint32_t nTris,
                                                uint32_t* hist)
{
    __shared__ uint32_t sh[HXR_KDB_HIST];
    const uint32_t base = blockIdx.x * HXR_KDB_CHUNK;
    if (base >= nRefs) return;
    const uint32_t first = refNode[base];
    for (uint32_t t = threadIdx.x; t < HXR_KDB_HIST; t += blockDim.x) sh[t] = 0;
    __syncthreads();
    for (uint32_t k = 0; k < HXR_KDB_CHUNK / 256u; k++) {
        const uint32_t i = base + k * 256u + threadIdx.x;
        if (i >= nRefs) break;
        const uint32_t node = refNode[i];
        const kdb::NodeWork* w = work + node;
        if ((int)w->count <= P.binnedAbove) continue;
        const uint32_t t = refTri[i];
        for (int a = 0; a < 3; a++) {
            const double bmn = w->mn[a], bmx = w->mx[a];
            if (!(bmx - bmn > 0)) continue;
            int b0, b1;
            kdb::bin_range(tb[(size_t)a * nTris + t], tb[(size_t)(3 + a) * nTris + t], bmn, bmx, b0, b1);
            if (node == first) {
                atomicAdd(&sh[(a * 2) * HXR_KDB_BINS + b0], 1u);
                atomicAdd(&sh[(a * 2 + 1) * HXR_KDB_BINS + b1], 1u);
            } else {
                atomicAdd(hist + (size_t)node * HXR_KDB_HIST + (a * 2) * HXR_KDB_BINS + b0, 1u);
                atomicAdd(hist + (size_t)node * HXR_KDB_HIST + (a * 2 + 1) * HXR_KDB_BINS + b1, 1u);
            }
        }
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < HXR_KDB_HIST; t += blockDim.x)
        if (sh[t]) atomicAdd(hist + (size_t)first * HXR_KDB_HIST + t, sh[t]);
}

__global__ void __launch_bounds__(128) k_kd_choose(kdb::Params P, const kdb::NodeWork* __restrict__ work, uint32_t nNodes, int depth,
                                                   const uint32_t* __restrict__ hist, const uint32_t* __restrict__ refTri, const double* __restrict__ tb,
                                                   uint32_t nTris, kdb::Decision* dec)
{
    __shared__ double smn[4][HXR_KDB_EXACT_MAX], smx[4][HXR_KDB_EXACT_MAX];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t warps = gridDim.x * 4u;
    for (uint32_t node = blockIdx.x * 4u + wib; node < nNodes; node += warps) {
        const kdb::NodeWork w = work[node];
        const uint32_t n = w.count;
        kdb::Decision d;
        d.split = 0; d.axis = 3; d.bad = w.bad; d.nl = 0;
        if (n > 1 && depth < P.maxDepth) {
            float bestCost = INFINITY, bestSplit = 0;
            int bestAxis = -1;
            for (int axis = 0; axis < 3; axis++) {
                if (!(w.mx[axis] - w.mn[axis] > 0)) continue;  // (warp-uniform)
                float ac = INFINITY, as = 0;
                if ((int)n > P.binnedAbove) {
                    // lane k owns plane k: the references that START in bins below it lie (at least partly) on its left,
                    // those that END in bins below it do not reach its right
                    const uint32_t* sc = hist + ((size_t)node * 3 + axis) * 2 * HXR_KDB_BINS;
                    const uint32_t s = sc[lane], e = sc[HXR_KDB_BINS + lane];
                    uint32_t ps = s, pe = e;
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t a = __shfl_up_sync(full, ps, o), b = __shfl_up_sync(full, pe, o);
                        if (lane >= o) { ps += a; pe += b; }
                    }
                    if (lane >= 1) {
                        const float sp = kdb::binned_plane(w, axis, lane);
                        if (kdb::plane_inside(w, axis, sp)) { ac = kdb::sah_cost(P, w, axis, sp, ps - s, n - (pe - e)); as = sp; }
                    }
                } else {
                    // small node: its references' bounds on this axis go to shared memory, every bound edge is a candidate plane
                    __syncwarp();
                    for (uint32_t j = lane; j < n; j += 32) {
                        const uint32_t t = refTri[w.start + j];
                        smn[wib][j] = tb[(size_t)axis * nTris + t];
                        smx[wib][j] = tb[(size_t)(3 + axis) * nTris + t];
                    }
                    __syncwarp();
                    for (uint32_t ci = lane; ci < 2 * n; ci += 32) {
                        const float sp = (float)(ci < n ? smn[wib][ci] : smx[wib][ci - n]);
                        if (!kdb::plane_inside(w, axis, sp)) continue;
                        const double sd = (double)sp;
                        uint32_t nl = 0, nr = 0;
                        for (uint32_t j = 0; j < n; j++) {
                            const double mn = smn[wib][j], mx = smx[wib][j];
                            nl += (mn < sd || (mn == sd && mx == sd)) ? 1u : 0u;
                            nr += mx > sd ? 1u : 0u;
                        }
                        const float c = kdb::sah_cost(P, w, axis, sd, nl, nr);
                        if (kdb::better(c, sp, ac, as)) { ac = c; as = sp; }
                    }
                }
                for (int o = 16; o; o >>= 1) {
                    const float oc = __shfl_xor_sync(full, ac, o), os = __shfl_xor_sync(full, as, o);
                    if (kdb::better(oc, os, ac, as)) { ac = oc; as = os; }
                }
                if (ac < bestCost) { bestCost = ac; bestAxis = axis; bestSplit = as; }
            }
            if (bestAxis >= 0) {
                uint32_t bad = w.bad;
                if (kdb::keep_split(P, n, bestCost, bad)) { d.split = bestSplit; d.axis = bestAxis; d.bad = bad; }
            }
        }
        if (lane == 0) dec[node] = d;
    }
}

__global__ void __launch_bounds__(256) k_kd_classify(const uint32_t* __restrict__ refTri, const uint32_t* __restrict__ refNode, uint32_t nRefs,
                                                     const kdb::Decision* __restrict__ dec, const double* __restrict__ tb, uint32_t nTris, uint32_t* flagL,
                                                     uint32_t* flagR)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nRefs; i += stride) kdb::classify_item(i, refTri, refNode, dec, tb, nTris, flagL, flagR);
    if (blockIdx.x == 0 && threadIdx.x == 0) flagL[nRefs] = flagR[nRefs] = 0;
}

__global__ void __launch_bounds__(256) k_kd_plan(const kdb::NodeWork* __restrict__ work, uint32_t nNodes, kdb::Decision* dec, const uint32_t* __restrict__ scanL,
                                                 const uint32_t* __restrict__ scanR, uint32_t* childRefs, uint32_t* isSplit, uint32_t* leafRefs, uint32_t* levelMax)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t big = 0;
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < nNodes; n += stride) {
        kdb::plan_item(n, work, dec, scanL, scanR, childRefs, isSplit, leafRefs);
        if (isSplit[n]) big = max(big, max(dec[n].nl, childRefs[n] - dec[n].nl));
    }
    if (big) atomicMax(levelMax, big);
    if (blockIdx.x == 0 && threadIdx.x == 0) childRefs[nNodes] = isSplit[nNodes] = leafRefs[nNodes] = 0;
}

__global__ void __launch_bounds__(256) k_kd_emit(const kdb::NodeWork* __restrict__ work, uint32_t nNodes, const kdb::Decision* __restrict__ dec,
                                                 const uint32_t* __restrict__ childRefs, const uint32_t* __restrict__ isSplit,
                                                 const uint32_t* __restrict__ leafRefs, uint32_t outCount, uint32_t leafBase, kdb::OutNode* out, kdb::NodeWork* next)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < nNodes; n += stride)
        kdb::emit_item(n, work, dec, childRefs, isSplit, leafRefs, outCount, leafBase, out, next);
}

__global__ void __launch_bounds__(256) k_kd_scatter(const uint32_t* __restrict__ refTri, const uint32_t* __restrict__ refNode, uint32_t nRefs,
                                                    const kdb::NodeWork* __restrict__ work, const kdb::Decision* __restrict__ dec,
                                                    const uint32_t* __restrict__ scanL, const uint32_t* __restrict__ scanR,
                                                    const uint32_t* __restrict__ childRefs, const uint32_t* __restrict__ isSplit,
                                                    const uint32_t* __restrict__ leafRefs, uint32_t leafBase, uint32_t* nextTri, uint32_t* nextNode,
                                                    uint32_t* leafOut)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nRefs; i += stride)
        kdb::scatter_item(i, refTri, refNode, work, dec, scanL, scanR, childRefs, isSplit, leafRefs, leafBase, nextTri, nextNode, leafOut);
}

static uint32_t kd_grid(const Context* c, uint32_t n, uint32_t perBlock = 256)
{
    return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)n + perBlock - 1) / perBlock, (uint64_t)c->sms * 16));
}

int kd_bounds(Context* c, const double* vertices, const int32_t* triV, uint32_t nTris, double* tb)
{
    LaunchScope ls(c, PROF_OTHER);
    k_kd_bounds<<<kd_grid(c, nTris), 256, 0, c->stream>>>(vertices, triV, nTris, tb);
    return 1;
}
int kd_iota(Context* c, uint32_t* refTri, uint32_t* refNode, uint32_t n)
{
    LaunchScope ls(c, PROF_OTHER);
    k_kd_iota<<<kd_grid(c, n), 256, 0, c->stream>>>(refTri, refNode, n);
    return 1;
}
int kd_bin(Context* c, const kdb::Params& P, const kdb::NodeWork* work, const uint32_t* refTri, const uint32_t* refNode, uint32_t nRefs, const double* tb,
           uint32_t nTris, uint32_t* hist)
{
    if (!nRefs) return 0;
    LaunchScope ls(c, PROF_OTHER);
    k_kd_bin<<<(nRefs + HXR_KDB_CHUNK - 1) / HXR_KDB_CHUNK, 256, 0, c->stream>>>(P, work, refTri, refNode, nRefs, tb, nTris, hist);
    return 1;
}
int kd_choose(Context* c, const kdb::Params& P, const kdb::NodeWork* work, uint32_t nNodes, int depth, const uint32_t* hist, const uint32_t* refTri,
              const double* tb, uint32_t nTris, kdb::Decision* dec)
{
    if (!nNodes) return 0;
    LaunchScope ls(c, PROF_OTHER);
    const uint32_t blocks = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)nNodes + 3) / 4, (uint64_t)c->sms * 16));
    k_kd_choose<<<blocks, 128, 0, c->stream>>>(P, work, nNodes, depth, hist, refTri, tb, nTris, dec);
    return 1;
}
int kd_classify(Context* c, const uint32_t* refTri, const uint32_t* refNode, uint32_t nRefs, const kdb::Decision* dec, const double* tb, uint32_t nTris,
                uint32_t* flagL, uint32_t* flagR)
{
    LaunchScope ls(c, PROF_OTHER);
    k_kd_classify<<<kd_grid(c, nRefs), 256, 0, c->stream>>>(refTri, refNode, nRefs, dec, tb, nTris, flagL, flagR);
    return 1;
}
int scan_u32(Context* c, uint32_t* data, uint32_t n)
{
    if (!n) return 0;
    LaunchScope ls(c, PROF_OTHER);
    size_t need = 0;
    if (!ck(c, cub::DeviceScan::ExclusiveSum(nullptr, need, data, data, (int)n, c->stream), "scan (size query)")) return 0;
    if (need > c->scanTmpBytes) {
        if (c->scanTmp) { cudaStreamSynchronize(c->stream); cudaFree(c->scanTmp); }
        c->scanTmp = nullptr;
        c->scanTmpBytes = 0;
        if (!ck(c, cudaMalloc(&c->scanTmp, need + need / 2), "scan scratch allocation")) return 0;
        c->scanTmpBytes = need + need / 2;
    }
    size_t have = c->scanTmpBytes;
    ck(c, cub::DeviceScan::ExclusiveSum(c->scanTmp, have, data, data, (int)n, c->stream), "scan");
    return 1;
}
int kd_plan(Context* c, const kdb::NodeWork* work, uint32_t nNodes, kdb::Decision* dec, const uint32_t* scanL, const uint32_t* scanR, uint32_t* childRefs,
            uint32_t* isSplit, uint32_t* leafRefs, uint32_t* levelMax)
{
    LaunchScope ls(c, PROF_OTHER);
    k_kd_plan<<<kd_grid(c, nNodes), 256, 0, c->stream>>>(work, nNodes, dec, scanL, scanR, childRefs, isSplit, leafRefs, levelMax);
    return 1;
}
int kd_emit(Context* c, const kdb::NodeWork* work, uint32_t nNodes, const kdb::Decision* dec, const uint32_t* childRefs, const uint32_t* isSplit,
            const uint32_t* leafRefs, uint32_t outCount, uint32_t leafBase, kdb::OutNode* out, kdb::NodeWork* next)
{
    LaunchScope ls(c, PROF_OTHER);
    k_kd_emit<<<kd_grid(c, nNodes), 256, 0, c->stream>>>(work, nNodes, dec, childRefs, isSplit, leafRefs, outCount, leafBase, out, next);
    return 1;
}
int kd_scatter(Context* c, const uint32_t* refTri, const uint32_t* refNode, uint32_t nRefs, const kdb::NodeWork* work, const kdb::Decision* dec,
               const uint32_t* scanL, const uint32_t* scanR, const uint32_t* childRefs, const uint32_t* isSplit, const uint32_t* leafRefs,
               uint32_t leafBase, uint32_t* nextTri, uint32_t* nextNode, uint32_t* leafOut)
{
    if (!nRefs) return 0;
    LaunchScope ls(c, PROF_OTHER);
    k_kd_scatter<<<kd_grid(c, nRefs), 256, 0, c->stream>>>(refTri, refNode, nRefs, work, dec, scanL, scanR, childRefs, isSplit, leafRefs, leafBase, nextTri,
                                                           nextNode, leafOut);
    return 1;
}
