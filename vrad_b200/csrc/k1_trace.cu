// k1_trace.cu -- K1: batched kd-tree ray casting (closest hit) and segment visibility.
//
// Replaces Environment.Trace4Rays (raytracer/environment.go:140-145, a stub in the reference) for
// whole batches, and the per-lane post-processing of trace.TestLineDoesHitSky
// (raytracer/trace/testline.go:22-51).  One thread per ray; rays are SoA so every load and
// store of a warp is one coalesced 128-byte transaction; scene records come in through the
// read-only path as 64/128-bit vectors (common.cuh).  Algorithmic HBM bytes: 32 B/ray
// (closest hit: 24 in + 8 out), 24.125 B/segment (visibility).
#include "env_internal.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cmath>
#include <cstdlib>

namespace vrad {

constexpr int kTraceBlock = 128;
constexpr int kTraceWarps = kTraceBlock / 32;
constexpr int kRaysPerWarp = 512;              // each warp streams a contiguous chunk of the batch

// Closest hit (batched Trace4Rays).  Same supply as the visibility kernel below: persistent warps draw ranges of kTraceRange rays from
// a global counter and refill their lanes across ranges (round 1: a fixed 512-ray chunk per warp, drained at its end).
constexpr int kTraceRange = 128;
constexpr int kTraceBlocksPerSM = 10;
__global__ void __launch_bounds__(kTraceBlock, kTraceBlocksPerSM)
k1_trace_rays(DevScene S, int64_t n, const float* __restrict__ ox, const float* __restrict__ oy,
              const float* __restrict__ oz, const float* __restrict__ dx, const float* __restrict__ dy,
              const float* __restrict__ dz, const float* __restrict__ tmin, const float* __restrict__ tmax,
              int skip_id, int32_t* __restrict__ hit_tri, int32_t* __restrict__ hit_sid,
              float* __restrict__ hit_t, float* __restrict__ normal_soa, unsigned long long* __restrict__ counter) {
    auto fetch = [&](int64_t& i, Ray& r, float& t0, float& t1, float& len) {
        r = Ray{__ldcs(&ox[i]), __ldcs(&oy[i]), __ldcs(&oz[i]), __ldcs(&dx[i]), __ldcs(&dy[i]), __ldcs(&dz[i])};   // streamed once
        t0 = tmin ? __ldcs(&tmin[i]) : 0.0f; t1 = __ldcs(&tmax[i]); len = 0.0f;
        return true;
    };
    auto retire = [&](int64_t i, int tri, float t, float) {
        if (hit_tri) hit_tri[i] = tri;
        if (hit_t) hit_t[i] = t;
        if (hit_sid) hit_sid[i] = tri >= 0 ? __float_as_int(__ldg(&S.q2[tri]).z) : -1;
        if (normal_soa) {
            const float4 q = tri >= 0 ? __ldg(&S.q0[tri]) : make_float4(0.f, 0.f, 0.f, 0.f);
            normal_soa[i] = q.x; normal_soa[n + i] = q.y; normal_soa[2 * n + i] = q.z;
        }
    };
    auto more = [&](int64_t& next, int64_t& end) {
        unsigned long long c = 0;
        if ((threadIdx.x & 31) == 0) c = atomicAdd(counter, (unsigned long long)kTraceRange);
        c = __shfl_sync(0xffffffffu, c, 0);
        if ((int64_t)c >= n) return false;
        next = (int64_t)c; end = (int64_t)c + kTraceRange < n ? (int64_t)c + kTraceRange : n;
        return true;
    };
    stream_rays<decltype(fetch), decltype(retire), false, false, decltype(more)>(S, 0, 0, skip_id, fetch, retire, nullptr, more);
}

template <bool SKY, bool TOP>
__global__ void __launch_bounds__(kTraceBlock, TOP ? 8 : 10)
k1_test_lines(DevScene S, int64_t n, int64_t stride, const float* __restrict__ a, const float* __restrict__ b,
              uint32_t* __restrict__ bits, int rpw) {   // a, b: SoA blocks x[stride] y[stride] z[stride]; rpw = rays per warp chunk (multiple of 32, <= kRaysPerWarp)
    __shared__ uint32_t words[kTraceWarps][kRaysPerWarp / 32];
    extern __shared__ int2 top_s[];
    if (TOP) {
        for (int i = threadIdx.x; i < S.n_top; i += blockDim.x) top_s[i] = __ldg(&S.top[i]);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_chunks = (n + rpw - 1) / rpw;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t chunk = warp0; chunk < n_chunks; chunk += nwarps) {
        const int64_t base = chunk * rpw;
        const int64_t end = base + rpw < n ? base + rpw : n;
        if (lane < kRaysPerWarp / 32) words[warp][lane] = 0u;
        __syncwarp();
        auto fetch = [&](int64_t& i, Ray& r, float& t0, float& t1, float& len) {
            t0 = 0.0f;
            r = Ray{0.f, 0.f, 0.f, 1.f, 1.f, 1.f}; len = 0.0f;
            const bool ok = segment_to_ray(__ldcs(&a[i]), __ldcs(&a[stride + i]), __ldcs(&a[2 * stride + i]),
                                           __ldcs(&b[i]), __ldcs(&b[stride + i]), __ldcs(&b[2 * stride + i]), r, len);   // streamed once
            t1 = len;
            return ok;
        };
        auto retire = [&](int64_t i, int tri, float t, float len) {
            // occlusion rule of raytracer/trace/testline.go:42-51
            bool occluded = tri != -1 && t < len;
            if (SKY && occluded) occluded = (__float_as_int(__ldg(&S.q2[tri]).z) & 0x01000000) == 0;
            if (!occluded) atomicOr(&words[warp][(int)(i - base) >> 5], 1u << ((int)(i - base) & 31));
        };
        stream_rays<decltype(fetch), decltype(retire), !SKY, TOP>(S, base, end, -1, fetch, retire, top_s);
        __syncwarp();
        const int nw = (int)((end - base + 31) >> 5);
        if (lane < nw) bits[(base >> 5) + lane] = words[warp][lane];
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// Coherence pre-pass (r01 verdict: 28 % SIMT efficiency on random segments, L1 hit 36 % on the
// 1 M-triangle map).  The batch is traced in the order of a 30-bit key -- Morton code of the start point's cell
// (32^3 grid over the scene box), direction octant, Morton code of the end point's cell (16^3) -- so the 32 rays a
// warp holds at any time start next to each other and head the same way: they take the same branches of the
// tree, test the same leaves, and touch the same cache lines.  k1_sort_keys reads the segments once (coalesced),
// writes one 32-byte record per segment plus its key; cub's radix sort orders (key, index) pairs; the traversal
// kernel walks the sorted indices and fetches each record with one 32-byte sector read.  Results are per segment,
// so the order is invisible in the output (bit i of vis_bits is still segment i; set with a global atomicOr).
// Extra HBM traffic per segment: 24 read + 40 written by the key pass, ~64 moved by the sort, 36 read by the
// traversal = ~165 B against the 24.125 B the unsorted kernel streams -- about 0.45 ms per 2^24 segments.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread3(uint32_t v) {        // 000a bcde -> a00b00c00d00e
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ uint32_t cell_of(float x, float lo, float inv, int cells) {
    const float f = (x - lo) * inv;
    int c = (int)fminf(fmaxf(f, 0.0f), (float)(cells - 1));     // NaN -> 0 (fmaxf returns the non-NaN operand)
    return (uint32_t)c;
}

// key layouts (EnvOptions::k1_key), all 30 bits:
//   0  Morton(start cell, 32 per axis) | direction octant | Morton(end cell, 16 per axis)          -- start-major
//   1  6-D Morton: bits of start x,y,z and end x,y,z interleaved, 5 per coordinate                  -- both ends refine together
//   2  start cell on a grid of roughly CUBIC cells (up to 128 x 128 x 16, counts from the scene box) | octant | Morton(end, 8 per axis)
struct SortGrid { float lo[3], inv32[3], inv16[3], invc[3], inv8[3]; int nc[3]; int mode; };
__device__ __forceinline__ uint32_t spread6(uint32_t v) {        // 5 bits -> every 6th bit
    v &= 0x1fu;
    return (v & 1u) | ((v & 2u) << 5) | ((v & 4u) << 10) | ((v & 8u) << 15) | ((v & 16u) << 20);
}
__device__ __forceinline__ uint32_t spread2(uint32_t v) {        // 7 bits -> every 2nd bit
    v &= 0x7fu;
    v = (v | (v << 4)) & 0x0f0fu;
    v = (v | (v << 2)) & 0x3333u;
    v = (v | (v << 1)) & 0x5555u;
    return v;
}

// segment i of a batch: coordinates (SoA blocks, stride) or a pair of indices into the resident point table
struct SegSource {
    const float* a; const float* b; int64_t stride;
    const int2* pairs; const float4* pts; int n_pts;
};
template <bool INDEXED>
__device__ __forceinline__ void load_segment(const SegSource& src, int64_t i, float& ax, float& ay, float& az, float& bx, float& by, float& bz) {
    if (INDEXED) {
        int2 pr = __ldcs(&src.pairs[i]);
        pr.x = min(max(pr.x, 0), src.n_pts - 1); pr.y = min(max(pr.y, 0), src.n_pts - 1);    // memory safety; validity is checked by the entry point
        const float4 pa = __ldg(&src.pts[pr.x]), pb = __ldg(&src.pts[pr.y]);
        ax = pa.x; ay = pa.y; az = pa.z; bx = pb.x; by = pb.y; bz = pb.z;
    } else {
        ax = __ldcs(&src.a[i]); ay = __ldcs(&src.a[src.stride + i]); az = __ldcs(&src.a[2 * src.stride + i]);
        bx = __ldcs(&src.b[i]); by = __ldcs(&src.b[src.stride + i]); bz = __ldcs(&src.b[2 * src.stride + i]);
    }
}

template <bool INDEXED>
__global__ void __launch_bounds__(256)
k1_sort_keys(int64_t n, SegSource src, SortGrid G, float4* __restrict__ rec, uint32_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float ax, ay, az, bx, by, bz;
    load_segment<INDEXED>(src, i, ax, ay, az, bx, by, bz);
    if (!INDEXED) {      // index pairs are small enough (8 B, one sector) to be fetched again through the order; coordinates get a 32-byte record
        rec[2 * i] = make_float4(ax, ay, az, bx);
        rec[2 * i + 1] = make_float4(by, bz, 0.f, 0.f);
    }
    const uint32_t oct = (bx < ax ? 1u : 0u) | (by < ay ? 2u : 0u) | (bz < az ? 4u : 0u);
    uint32_t key;
    if (G.mode == 1) {
        key = (spread6(cell_of(ax, G.lo[0], G.inv32[0], 32)) << 5) | (spread6(cell_of(ay, G.lo[1], G.inv32[1], 32)) << 4) |
              (spread6(cell_of(az, G.lo[2], G.inv32[2], 32)) << 3) | (spread6(cell_of(bx, G.lo[0], G.inv32[0], 32)) << 2) |
              (spread6(cell_of(by, G.lo[1], G.inv32[1], 32)) << 1) | spread6(cell_of(bz, G.lo[2], G.inv32[2], 32));
    } else if (G.mode == 2) {
        const uint32_t cxy = spread2(cell_of(ax, G.lo[0], G.invc[0], G.nc[0])) | (spread2(cell_of(ay, G.lo[1], G.invc[1], G.nc[1])) << 1);
        const uint32_t cz = cell_of(az, G.lo[2], G.invc[2], G.nc[2]);
        const uint32_t me = spread3(cell_of(bx, G.lo[0], G.inv8[0], 8)) | (spread3(cell_of(by, G.lo[1], G.inv8[1], 8)) << 1) |
                            (spread3(cell_of(bz, G.lo[2], G.inv8[2], 8)) << 2);
        key = (((cxy << 4) | cz) << 12) | (oct << 9) | me;
    } else {
        const uint32_t mo = spread3(cell_of(ax, G.lo[0], G.inv32[0], 32)) | (spread3(cell_of(ay, G.lo[1], G.inv32[1], 32)) << 1) |
                            (spread3(cell_of(az, G.lo[2], G.inv32[2], 32)) << 2);
        const uint32_t me = spread3(cell_of(bx, G.lo[0], G.inv16[0], 16)) | (spread3(cell_of(by, G.lo[1], G.inv16[1], 16)) << 1) |
                            (spread3(cell_of(bz, G.lo[2], G.inv16[2], 16)) << 2);
        key = (mo << 15) | (oct << 12) | me;
    }
    keys[i] = key;
    idx[i] = (uint32_t)i;
}

// traversal over the sorted order: position j of the order is segment perm[j].  Persistent warps draw ranges of
// kSortedRange positions from a global counter: neighbouring positions cost alike (that is the point of the order), so
// fixed per-warp chunks would leave the kernel waiting for the warps that drew the expensive corner of the scene
// (r02, first form: 512-position chunks, S3 map, 2^22 segments: 7.2 ms sorted against 5.2 ms unsorted).
// one thread waits until a chunk of a host batch has landed (see `arrive` below); out of line: its registers are not the traversal's
__device__ __noinline__ void wait_for_chunk(const uint32_t* flag, uint32_t* err) {
    const long long t0 = clock64();
    uint32_t v;
    do {
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (!v && clock64() - t0 > (1LL << 31)) { atomicExch(err, 1u); break; }
    } while (!v);
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
}

constexpr int kSortedRange = 128;
constexpr int kSortedBlocksPerSM = 12;       // 40 registers: 48 resident warps (the traversal waits on memory: on the 1 M-triangle map L1 hits are 44 %)
template <bool SKY, bool TOP, bool INDEXED, int BPSM = kSortedBlocksPerSM>
__global__ void __launch_bounds__(kTraceBlock, TOP ? 8 : BPSM)
k1_test_lines_sorted(DevScene S, int64_t n, const uint32_t* __restrict__ perm, const float4* __restrict__ rec, SegSource src, uint32_t* __restrict__ bits,
                     unsigned long long* __restrict__ counter, const uint32_t* arrive, int chunk_shift, uint32_t* arrive_err) {
    extern __shared__ int2 top_s[];
    if (TOP) {
        for (int i = threadIdx.x; i < S.n_top; i += blockDim.x) top_s[i] = __ldg(&S.top[i]);
        __syncthreads();
    }
    // the lane's slot in `my_ray` carries the ORIGINAL segment number once fetched
    auto fetch = [&](int64_t& i, Ray& r, float& t0, float& t1, float& len) {
        const uint32_t seg = perm ? __ldcs(&perm[i]) : (uint32_t)i;        // no order given: the batch as it came, same streaming
        float ax, ay, az, bx, by, bz;
        if (INDEXED) load_segment<true>(src, (int64_t)seg, ax, ay, az, bx, by, bz);
        else if (!rec) load_segment<false>(src, (int64_t)seg, ax, ay, az, bx, by, bz);
        else {
            const float4 q0 = __ldcs(&rec[2 * (int64_t)seg]), q1 = __ldcs(&rec[2 * (int64_t)seg + 1]);
            ax = q0.x; ay = q0.y; az = q0.z; bx = q0.w; by = q1.x; bz = q1.y;
        }
        i = seg;
        t0 = 0.0f;
        r = Ray{0.f, 0.f, 0.f, 1.f, 1.f, 1.f}; len = 0.0f;
        const bool ok = segment_to_ray(ax, ay, az, bx, by, bz, r, len);
        t1 = len;
        return ok;
    };
    auto retire = [&](int64_t i, int tri, float t, float len) {
        // occlusion rule of raytracer/trace/testline.go:42-51
        bool occluded = tri != -1 && t < len;
        if (SKY && occluded) occluded = (__float_as_int(__ldg(&S.q2[tri]).z) & 0x01000000) == 0;
        if (!occluded) atomicOr(&bits[i >> 5], 1u << ((int)i & 31));
    };
    // `arrive` (host batches): the batch is still on its way over PCIe while this kernel runs -- chunk k of 2^chunk_shift
    // positions may be read once arrive[k] is set (a 4-byte copy queued behind the chunk's data on the copy stream; chunks land in
    // order).  ONE launch serves the whole batch: no per-chunk kernel, no per-chunk tail.  arrive[n_chunks] is the error word: a
    // wait of ~2^31 cycles gives up (the copy stream died) and the entry point reports it.
    __shared__ int ready_s[kTraceWarps];          // highest chunk this warp has seen arrive (kept out of the traversal's registers)
    if ((threadIdx.x & 31) == 0) ready_s[threadIdx.x >> 5] = -1;
    __syncwarp();
    auto more = [&](int64_t& next, int64_t& end) {
        unsigned long long c = 0;
        if ((threadIdx.x & 31) == 0) c = atomicAdd(counter, (unsigned long long)kSortedRange);
        c = __shfl_sync(0xffffffffu, c, 0);
        if ((int64_t)c >= n) return false;
        if (arrive) {
            const int ch = (int)(c >> chunk_shift);
            if (ch > ready_s[threadIdx.x >> 5]) {
                __syncwarp();
                if ((threadIdx.x & 31) == 0) { wait_for_chunk(arrive + ch, arrive_err); ready_s[threadIdx.x >> 5] = ch; }
                __syncwarp();
            }
        }
        next = (int64_t)c; end = (int64_t)c + kSortedRange < n ? (int64_t)c + kSortedRange : n;
        return true;
    };
    stream_rays<decltype(fetch), decltype(retire), !SKY, TOP, decltype(more)>(S, 0, 0, -1, fetch, retire, top_s, more);
}

// unsorted traversal of index pairs (the coordinate form's k1_test_lines with the endpoints fetched from the point table)
template <bool SKY>
__global__ void __launch_bounds__(kTraceBlock)
k1_test_lines_indexed(DevScene S, int64_t n, SegSource src, uint32_t* __restrict__ bits, int rpw) {
    __shared__ uint32_t words[kTraceWarps][kRaysPerWarp / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_chunks = (n + rpw - 1) / rpw;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t chunk = warp0; chunk < n_chunks; chunk += nwarps) {
        const int64_t base = chunk * rpw;
        const int64_t end = base + rpw < n ? base + rpw : n;
        if (lane < kRaysPerWarp / 32) words[warp][lane] = 0u;
        __syncwarp();
        auto fetch = [&](int64_t& i, Ray& r, float& t0, float& t1, float& len) {
            float ax, ay, az, bx, by, bz;
            load_segment<true>(src, i, ax, ay, az, bx, by, bz);
            t0 = 0.0f;
            r = Ray{0.f, 0.f, 0.f, 1.f, 1.f, 1.f}; len = 0.0f;
            const bool ok = segment_to_ray(ax, ay, az, bx, by, bz, r, len);
            t1 = len;
            return ok;
        };
        auto retire = [&](int64_t i, int tri, float t, float len) {
            bool occluded = tri != -1 && t < len;
            if (SKY && occluded) occluded = (__float_as_int(__ldg(&S.q2[tri]).z) & 0x01000000) == 0;
            if (!occluded) atomicOr(&words[warp][(int)(i - base) >> 5], 1u << ((int)(i - base) & 31));
        };
        stream_rays<decltype(fetch), decltype(retire), !SKY>(S, base, end, -1, fetch, retire);
        __syncwarp();
        const int nw = (int)((end - base + 31) >> 5);
        if (lane < nw) bits[(base >> 5) + lane] = words[warp][lane];
        __syncwarp();
    }
}

// one warp per kRaysPerWarp-ray chunk, capped at 64 resident-block waves per SM (grid-stride beyond that)
// rays per warp chunk: 512 for big batches; smaller batches take shorter chunks so that there are still about four chunks per
// resident warp (a 2^21-segment batch in 512-ray chunks is 4096 warps for 5920 warp slots: r02, the chunked host path ran its
// kernels 30 % below the rate of one 2^24 batch)
static int rays_per_warp(const vrad_env* e, int64_t n) {
    const int64_t want = n / ((int64_t)e->sm_count * 40 * 4);
    return (int)std::min<int64_t>(kRaysPerWarp, std::max<int64_t>(64, want & ~(int64_t)31));
}
static int stream_grid(const vrad_env* e, int64_t n, int rpw = kRaysPerWarp) {
    int64_t blocks = ((n + rpw - 1) / rpw + kTraceWarps - 1) / kTraceWarps;
    const int64_t cap = (int64_t)e->sm_count * 64;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int launch_trace_rays(vrad_env* e, int64_t n, const float* ox, const float* oy, const float* oz, const float* dx,
                      const float* dy, const float* dz, const float* tmin, const float* tmax, int32_t skip_id,
                      int32_t* hit_tri, int32_t* hit_sid, float* hit_t, float* normal_soa) {
    void* d_ctr;
    int rc = scratch_get(e, 19, 8, &d_ctr);
    if (rc) return rc;
    timing_begin(e);
    VRAD_CUDA_CHECK(cudaMemsetAsync(d_ctr, 0, 8, e->stream));
    const int grid = (int)std::min<int64_t>((int64_t)e->sm_count * kTraceBlocksPerSM, (n + kTraceRange * kTraceWarps - 1) / (kTraceRange * kTraceWarps));
    k1_trace_rays<<<std::max(grid, 1), kTraceBlock, 0, e->stream>>>(
        e->scene, n, ox, oy, oz, dx, dy, dz, tmin, tmax, skip_id, hit_tri, hit_sid, hit_t, normal_soa, (unsigned long long*)d_ctr);
    timing_end(e, 1);
    VRAD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// Order policy: VRAD_K1_SORT=0 never, 1 always, unset = batches of at least kSortMin segments.
constexpr int64_t kSortMin = (int64_t)1 << 16;
constexpr int64_t kSortBatch = (int64_t)1 << 24;          // segments ordered at a time (bounds the scratch: 48 B per segment)
// Automatic rule: order big batches on scenes whose tree and triangles fit one SM's L1 (256 KB).  There the traversal is purely
// issue-bound and ordering halves the warp instructions (r02, S1 box room, 58 KB: 15.0 instead of 8.7 threads per instruction,
// 2.66 ms against 3.23 ms per 2^24 segments, sort included).  It stops paying as soon as the scene spills out of L1: on the
// 49,586-triangle map (3 MB) 2.95 ms ordered against 2.87 ms unordered; on the 1 M-triangle map (56 MB) the traversal waits on
// memory, the ordered warps still diverge in the leaves (8.0 threads per instruction either way) and the sort is not recovered:
// 4.28 ms ordered against 3.87 ms unordered per 2^22 segments.
constexpr size_t kSortSceneBytes = (size_t)256 << 10;
static bool want_sort(const vrad_env* e, int64_t n) {
    const int mode = e->opt.k1_sort;
    if (mode >= 0) return mode != 0;
    const size_t scene_bytes = (size_t)e->scene.n_nodes * 8 + (size_t)e->scene.n_tris * 48 + (size_t)e->scene.n_idx * 4;
    return n >= kSortMin && scene_bytes <= kSortSceneBytes;
}

static SortGrid sort_grid(const vrad_env* e) {
    SortGrid G;
    G.mode = e->opt.k1_key;
    if (G.mode < 0) {        // automatic: cells should be roughly cubic -- a flat world (outdoor map, 16:16:1) takes the cubic-cell layout
        float wmin = 1e30f, wmax = 0.0f;
        for (int c = 0; c < 3; c++) { const float w = e->scene.bmax[c] - e->scene.bmin[c]; wmin = std::min(wmin, w); wmax = std::max(wmax, w); }
        G.mode = wmax > 4.0f * std::max(wmin, 1.0f) ? 2 : 0;
    }
    double vol = 1.0;
    for (int c = 0; c < 3; c++) vol *= std::max(1.0, (double)e->scene.bmax[c] - (double)e->scene.bmin[c]);
    const double cell = std::cbrt(vol / 262144.0);           // edge of a cubic cell if 2^18 of them filled the box
    const int cap[3] = {128, 128, 16};
    for (int c = 0; c < 3; c++) {
        const float w = e->scene.bmax[c] - e->scene.bmin[c];
        G.lo[c] = e->scene.bmin[c];
        G.inv32[c] = w > 0.0f ? 32.0f / w : 0.0f;
        G.inv16[c] = w > 0.0f ? 16.0f / w : 0.0f;
        G.inv8[c] = w > 0.0f ? 8.0f / w : 0.0f;
        G.nc[c] = std::min(cap[c], std::max(1, (int)std::lround(w / cell)));
        G.invc[c] = w > 0.0f ? (float)G.nc[c] / w : 0.0f;
    }
    return G;
}

// Enqueues the traversal of n segments (coordinates or index pairs) on e->stream; `bits` may be written with atomics
// (sorted order) or whole words.  *launches += kernels enqueued.  n must start on a 32-segment boundary of the output.
static int enqueue_test_lines(vrad_env* e, int64_t n, const SegSource& src, int sky_mode, uint32_t* bits, int* launches, bool host_chunk = false,
                              const uint32_t* arrive = nullptr, int chunk_shift = 0, uint32_t* arrive_err = nullptr) {
    const bool indexed = src.pairs != nullptr;
    // chunks of a host batch are traced as they arrive unless ordering is forced: the call is bound by PCIe or close to it, and a
    // sort per 2^21-segment chunk costs more than it saves (r02: 5.7 ms ordered against 5.0 ms per 2^24 index pairs)
    const bool sorted = (host_chunk && e->opt.k1_sort < 0) ? false : want_sort(e, n);
    if (!sorted && !e->opt.k1_stream) {
        const int rpw = rays_per_warp(e, n), g = stream_grid(e, n, rpw);
        if (indexed) {
            if (sky_mode) k1_test_lines_indexed<true><<<g, kTraceBlock, 0, e->stream>>>(e->scene, n, src, bits, rpw);
            else k1_test_lines_indexed<false><<<g, kTraceBlock, 0, e->stream>>>(e->scene, n, src, bits, rpw);
        } else {
            const size_t sm = (size_t)e->scene.n_top * sizeof(int2);
            if (sm) {
                if (sky_mode) k1_test_lines<true, true><<<g, kTraceBlock, sm, e->stream>>>(e->scene, n, src.stride, src.a, src.b, bits, rpw);
                else k1_test_lines<false, true><<<g, kTraceBlock, sm, e->stream>>>(e->scene, n, src.stride, src.a, src.b, bits, rpw);
            } else {
                if (sky_mode) k1_test_lines<true, false><<<g, kTraceBlock, 0, e->stream>>>(e->scene, n, src.stride, src.a, src.b, bits, rpw);
                else k1_test_lines<false, false><<<g, kTraceBlock, 0, e->stream>>>(e->scene, n, src.stride, src.a, src.b, bits, rpw);
            }
        }
        (*launches)++;
        return 0;
    }
    const SortGrid G = sort_grid(e);
    const int sort_lo = 30 - std::min(30, std::max(8, e->opt.k1_sort_bits));     // only the key's leading bits take part in the sort (one radix pass per 8)
    for (int64_t c0 = 0; c0 < n; c0 += kSortBatch) {
        const int64_t m = std::min(kSortBatch, n - c0);
        void *d_rec = nullptr, *d_k0 = nullptr, *d_k1 = nullptr, *d_i0 = nullptr, *d_i1 = nullptr, *d_tmp = nullptr;
        int rc;
        size_t tmp_bytes = 0;
        if (sorted) {
            if ((rc = scratch_get(e, 12, indexed ? 32 : (size_t)m * 32, &d_rec)) || (rc = scratch_get(e, 13, (size_t)m * 4, &d_k0)) || (rc = scratch_get(e, 14, (size_t)m * 4, &d_k1)) ||
                (rc = scratch_get(e, 15, (size_t)m * 4, &d_i0)) || (rc = scratch_get(e, 16, (size_t)m * 4, &d_i1))) return rc;
            cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint32_t*)d_k0, (uint32_t*)d_k1, (const uint32_t*)d_i0, (uint32_t*)d_i1, (int)m, sort_lo, 30, e->stream);
            if ((rc = scratch_get(e, 17, tmp_bytes + 16, &d_tmp))) return rc;
        }
        SegSource sub = src;
        if (indexed) sub.pairs = src.pairs + c0; else { sub.a = src.a + c0; sub.b = src.b + c0; }
        uint32_t* out = bits + (c0 >> 5);
        void* d_ctr;
        if ((rc = scratch_get(e, 19, 8, &d_ctr))) return rc;
        VRAD_CUDA_CHECK(cudaMemsetAsync(d_ctr, 0, 8, e->stream));
        VRAD_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)((m + 31) >> 5) * 4, e->stream));
        const int bpsm = e->opt.k1_bpsm == 10 || e->opt.k1_bpsm == 8 ? e->opt.k1_bpsm : kSortedBlocksPerSM;
        const int sgrid = (int)std::min<int64_t>((int64_t)e->sm_count * bpsm, (m + kSortedRange * kTraceWarps - 1) / (kSortedRange * kTraceWarps));
        if (sorted) {
            const int kb = (int)((m + 255) / 256);
            if (indexed) k1_sort_keys<true><<<kb, 256, 0, e->stream>>>(m, sub, G, (float4*)d_rec, (uint32_t*)d_k0, (uint32_t*)d_i0);
            else k1_sort_keys<false><<<kb, 256, 0, e->stream>>>(m, sub, G, (float4*)d_rec, (uint32_t*)d_k0, (uint32_t*)d_i0);
            VRAD_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, (const uint32_t*)d_k0, (uint32_t*)d_k1, (const uint32_t*)d_i0, (uint32_t*)d_i1, (int)m, sort_lo, 30, e->stream));
        }
        const size_t sm = (size_t)e->scene.n_top * sizeof(int2);
        const uint32_t* pm = (const uint32_t*)d_i1; const float4* rc4 = indexed ? nullptr : (const float4*)d_rec; unsigned long long* ctr = (unsigned long long*)d_ctr;
#define VRAD_SORTED_B(SKY, TOP, IDX, B) k1_test_lines_sorted<SKY, TOP, IDX, B><<<sgrid, kTraceBlock, (TOP) ? sm : 0, e->stream>>>(e->scene, m, pm, rc4, sub, out, ctr, arrive ? arrive + (c0 >> chunk_shift) : nullptr, chunk_shift, arrive_err)
#define VRAD_SORTED(SKY, TOP, IDX) do { if (bpsm == 10) VRAD_SORTED_B(SKY, TOP, IDX, 10); else if (bpsm == 8) VRAD_SORTED_B(SKY, TOP, IDX, 8); else VRAD_SORTED_B(SKY, TOP, IDX, kSortedBlocksPerSM); } while (0)
        if (indexed) {
            if (sm) { if (sky_mode) VRAD_SORTED(true, true, true); else VRAD_SORTED(false, true, true); }
            else { if (sky_mode) VRAD_SORTED(true, false, true); else VRAD_SORTED(false, false, true); }
        } else {
            if (sm) { if (sky_mode) VRAD_SORTED(true, true, false); else VRAD_SORTED(false, true, false); }
            else { if (sky_mode) VRAD_SORTED(true, false, false); else VRAD_SORTED(false, false, false); }
        }
#undef VRAD_SORTED
#undef VRAD_SORTED_B
        *launches += sorted ? 2 + 6 : 1;       // keys + traversal + cub's histogram / scan / onesweep passes (4 digit passes of 8 bits)
    }
    return 0;
}

int launch_test_lines(vrad_env* e, int64_t n, const float* start_soa, const float* stop_soa, int sky_mode, uint32_t* bits) {
    timing_begin(e);
    int launches = 0;
    const SegSource src{start_soa, stop_soa, n, nullptr, nullptr, 0};
    int rc = enqueue_test_lines(e, n, src, sky_mode, bits, &launches);
    timing_end(e, launches);
    if (rc) return rc;
    VRAD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

__global__ void k1_count_bad_indices(int64_t n2, const int32_t* __restrict__ idx, int n_pts, int* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool b = i < n2 && (idx[i] < 0 || idx[i] >= n_pts);
    const unsigned m = __ballot_sync(0xffffffffu, b);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(bad, __popc(m));
}

// a streaming launch that gave up waiting for its input marks the batch as bad (the index-pair entry point reads the count back)
__global__ void k1_fold_arrival_error(const uint32_t* __restrict__ err, int* __restrict__ bad) { if (*err && bad) atomicAdd(bad, 1 << 30); }

// index pairs on the device: count the indices outside the point table (synchronises the stream)
int check_pairs_on_device(vrad_env* e, int64_t n, const int32_t* d_pairs, int* bad_out) {
    void* d_bad;
    int rc = scratch_get(e, 18, 4, &d_bad);
    if (rc) return rc;
    VRAD_CUDA_CHECK(cudaMemsetAsync(d_bad, 0, 4, e->stream));
    k1_count_bad_indices<<<(int)((2 * n + 255) / 256), 256, 0, e->stream>>>(2 * n, d_pairs, (int)e->n_points, (int*)d_bad);
    return read_bad_index_count(e, bad_out);
}

int read_bad_index_count(vrad_env* e, int* bad_out) {
    void* d_bad;
    int rc = scratch_get(e, 18, 4, &d_bad);
    if (rc) return rc;
    VRAD_CUDA_CHECK(cudaMemcpyAsync(bad_out, d_bad, 4, cudaMemcpyDeviceToHost, e->stream));
    VRAD_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    VRAD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_test_lines_indexed(vrad_env* e, int64_t n, const int32_t* pairs, int sky_mode, uint32_t* bits) {
    timing_begin(e);
    int launches = 0;
    const SegSource src{nullptr, nullptr, 0, (const int2*)pairs, e->d_points.p, (int)e->n_points};
    int rc = enqueue_test_lines(e, n, src, sky_mode, bits, &launches);
    timing_end(e, launches);
    if (rc) return rc;
    VRAD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// Host-buffer path: the segments are copied in chunks on a second stream into two staging buffers
// while the previous chunk is traced, so the call costs max(PCIe, kernel) instead of their sum.
// h_a / h_b are host SoA blocks x[n] y[n] z[n] (coordinates), or h_pairs the host index pairs; d_bits is the
// device result (n bits).
// host_stride: distance between the x, y and z blocks of h_a / h_b (= n unless the batch is a range of a larger one)
int launch_test_lines_pipelined(vrad_env* e, int64_t n, const float* h_a, const float* h_b, int64_t host_stride, const int32_t* h_pairs, int sky_mode, uint32_t* d_bits) {
    // A short head chunk (2^19 segments: the traversal starts after 4 MB of pairs / 12 MB of coordinates), then equal chunks of
    // 2^21.  Whichever side is slower -- PCIe at 24 B per segment, the kernel at 8 -- the call costs that side plus one chunk of
    // the other.  (r02: chunks doubling up to 2^23 were worse for both forms -- a copy twice as long as the kernel before it
    // stalls the kernel stream, and a copy-bound call ends with the whole last kernel exposed.)
    // Unordered batches up to 2^26 segments go through ONE launch of the streaming kernel, which follows the copy stream chunk by
    // chunk (2^19 segments each) through arrival flags -- see k1_test_lines_sorted.
    const bool stream_one_launch = e->opt.k1_stream && e->opt.k1_sort <= 0 && n <= ((int64_t)1 << 26) && e->h_one;
    if (stream_one_launch) {
        const int kShift = h_pairs ? 19 : 21;          // 4 MB of pairs, 6 x 8 MB of coordinates per chunk: big enough for the copy engine
        const int64_t chunk = (int64_t)1 << kShift, n_chunks = (n + chunk - 1) >> kShift;
        void *d_stage, *d_arrive, *d_bad = nullptr;
        int rc;
        if ((rc = scratch_get(e, 20, (size_t)(h_pairs ? 8 : 24) * (size_t)n, &d_stage)) || (rc = scratch_get(e, 21, (size_t)(n_chunks + 1) * 4, &d_arrive))) return rc;
        if ((rc = scratch_get(e, 18, 4, &d_bad))) return rc;
        timing_begin(e);
        int launches = 0;
        VRAD_CUDA_CHECK(cudaMemsetAsync(d_arrive, 0, (size_t)(n_chunks + 1) * 4, e->stream));
        VRAD_CUDA_CHECK(cudaMemsetAsync(d_bad, 0, 4, e->stream));
        // the copies may start once earlier work on the main stream is done with the staging buffer and the flags are zero
        VRAD_CUDA_CHECK(cudaEventRecord(e->ev_done[0], e->stream));
        VRAD_CUDA_CHECK(cudaStreamWaitEvent(e->copy_stream, e->ev_done[0], 0));
        float* st = (float*)d_stage;
        for (int64_t c = 0; c < n_chunks; c++) {
            const int64_t c0 = c << kShift, m = std::min(chunk, n - c0);
            if (h_pairs) VRAD_CUDA_CHECK(cudaMemcpyAsync((int32_t*)d_stage + 2 * c0, h_pairs + 2 * c0, (size_t)m * 8, cudaMemcpyHostToDevice, e->copy_stream));
            else for (int k = 0; k < 3; k++) {
                VRAD_CUDA_CHECK(cudaMemcpyAsync(st + k * n + c0, h_a + k * host_stride + c0, (size_t)m * 4, cudaMemcpyHostToDevice, e->copy_stream));
                VRAD_CUDA_CHECK(cudaMemcpyAsync(st + (3 + k) * n + c0, h_b + k * host_stride + c0, (size_t)m * 4, cudaMemcpyHostToDevice, e->copy_stream));
            }
            VRAD_CUDA_CHECK(cudaMemcpyAsync((uint32_t*)d_arrive + c, e->h_one, 4, cudaMemcpyHostToDevice, e->copy_stream));
        }
        VRAD_CUDA_CHECK(cudaEventRecord(e->ev_copied[0], e->copy_stream));
        SegSource src{};
        if (h_pairs) { src.pairs = (const int2*)d_stage; src.pts = e->d_points.p; src.n_pts = (int)e->n_points; }
        else { src.a = st; src.b = st + 3 * n; src.stride = n; }
        rc = enqueue_test_lines(e, n, src, sky_mode, d_bits, &launches, true, (const uint32_t*)d_arrive, kShift, (uint32_t*)d_arrive + n_chunks);
        if (rc) return rc;
        // nothing later on the main stream may touch the staging buffer before the last copy is in (it is, once the kernel is done;
        // the wait makes the stream order say so too)
        VRAD_CUDA_CHECK(cudaStreamWaitEvent(e->stream, e->ev_copied[0], 0));
        if (h_pairs) { k1_count_bad_indices<<<(int)((2 * n + 255) / 256), 256, 0, e->stream>>>(2 * n, (const int32_t*)d_stage, (int)e->n_points, (int*)d_bad); launches++; }
        k1_fold_arrival_error<<<1, 1, 0, e->stream>>>((const uint32_t*)d_arrive + n_chunks, (int*)d_bad);
        timing_end(e, launches);
        VRAD_CUDA_CHECK(cudaGetLastError());
        return 0;
    }
    constexpr int64_t kChunkMin = (int64_t)1 << 19, kChunk = (int64_t)1 << 21;
    const int64_t stage_cap = std::min<int64_t>(kChunk, std::max<int64_t>(kChunkMin, n));
    for (int s = 0; s < 2; s++)
        if (e->d_stage[s].alloc((size_t)(h_pairs ? 2 : 6) * stage_cap)) { set_error("out of device memory for staging"); return VRAD_E_NOMEM; }
    void* d_bad = nullptr;
    if (h_pairs) {
        int rcb = scratch_get(e, 18, 4, &d_bad);
        if (rcb) return rcb;
        VRAD_CUDA_CHECK(cudaMemsetAsync(d_bad, 0, 4, e->stream));
    }
    timing_begin(e);
    int launches = 0;
    // the copy stream must not overwrite staging that earlier work on the main stream may still read
    VRAD_CUDA_CHECK(cudaEventRecord(e->ev_done[0], e->stream));
    VRAD_CUDA_CHECK(cudaEventRecord(e->ev_done[1], e->stream));
    int c = 0;
    int64_t step = kChunkMin;
    for (int64_t c0 = 0; c0 < n; c++) {
        const int s = c & 1;
        const int64_t m = std::min(std::min(step, stage_cap), n - c0);
        step = kChunk;
        float* st = e->d_stage[s].p;
        VRAD_CUDA_CHECK(cudaStreamWaitEvent(e->copy_stream, e->ev_done[s], 0));
        SegSource src{};
        if (h_pairs) {
            VRAD_CUDA_CHECK(cudaMemcpyAsync(st, h_pairs + 2 * c0, (size_t)m * 8, cudaMemcpyHostToDevice, e->copy_stream));
            src.pairs = (const int2*)st; src.pts = e->d_points.p; src.n_pts = (int)e->n_points;
        } else {
            for (int k = 0; k < 3; k++) {
                VRAD_CUDA_CHECK(cudaMemcpyAsync(st + k * stage_cap, h_a + k * host_stride + c0, (size_t)m * 4, cudaMemcpyHostToDevice, e->copy_stream));
                VRAD_CUDA_CHECK(cudaMemcpyAsync(st + (3 + k) * stage_cap, h_b + k * host_stride + c0, (size_t)m * 4, cudaMemcpyHostToDevice, e->copy_stream));
            }
            src.a = st; src.b = st + 3 * stage_cap; src.stride = stage_cap;
        }
        VRAD_CUDA_CHECK(cudaEventRecord(e->ev_copied[s], e->copy_stream));
        VRAD_CUDA_CHECK(cudaStreamWaitEvent(e->stream, e->ev_copied[s], 0));
        if (h_pairs) { k1_count_bad_indices<<<(int)((2 * m + 255) / 256), 256, 0, e->stream>>>(2 * m, (const int32_t*)st, (int)e->n_points, (int*)d_bad); launches++; }
        int rc = enqueue_test_lines(e, m, src, sky_mode, d_bits + (c0 >> 5), &launches, true);
        if (rc) return rc;
        VRAD_CUDA_CHECK(cudaEventRecord(e->ev_done[s], e->stream));
        c0 += m;
    }
    timing_end(e, launches);
    VRAD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

} // namespace vrad
