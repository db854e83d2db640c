// mesh_vis.cu -- occlusion-culling prepass: which base-mesh faces own at least one pixel.
//
// Replaces the OpenGL hardware rasterisation behind nvdiffrast's dr.rasterize as Frosting uses it
// (frosting_utils/nvdiffrast.py:42-54, frosting_utils/mesh_rasterization.py:121-148):
//   clip = [v, 1] @ full_proj_transform;  GL clip volume -w <= z <= w;  pixel-centre sampling with a
//   top-left fill rule;  nearest depth wins (ties: lowest face id, i.e. first drawn under GL_LESS);
//   output face id per pixel (-1 = background), no back-face culling.
// On the hot path only the SET of visible faces is consumed (frosting_model.py:1534-1539,
// frosting_trainers/refine.py:436-441).
//
// Design: the shell base has ~1-2 M triangles at 1080p, i.e. most triangles cover a pixel or two, so
// the rasteriser is a per-triangle scan with a 64-bit atomicMin z-buffer (depth_bits<<32 | face):
// one thread per small triangle, one warp / one CTA per larger one (size-classed by bounding box,
// three launches of the same kernel), followed by a resolve pass.  No binning is needed at this
// triangle size and the z-buffer (16.6 MB at 1080p) lives in L2.
#include "common.cuh"

namespace fb200 {

namespace {

typedef unsigned long long u64;

struct ClipV { float x, y, z, w; };

// All arithmetic below uses explicit round-to-nearest intrinsics (no compiler-chosen FMA contraction) so
// that oracle/raster_oracle.c reproduces every coverage decision bit for bit.
__device__ __forceinline__ ClipV to_clip(const float* __restrict__ m, float x, float y, float z) {
    ClipV c;
    c.x = affine_row(m, 0, x, y, z);
    c.y = affine_row(m, 1, x, y, z);
    c.z = affine_row(m, 2, x, y, z);
    c.w = affine_row(m, 3, x, y, z);
    return c;
}

__device__ __forceinline__ float lerp1(float a, float b, float t) { return ffma(fadd(b, -a), t, a); }
__device__ __forceinline__ ClipV lerp(const ClipV& a, const ClipV& b, float t) {
    ClipV c;
    c.x = lerp1(a.x, b.x, t); c.y = lerp1(a.y, b.y, t); c.z = lerp1(a.z, b.z, t); c.w = lerp1(a.w, b.w, t);
    return c;
}

struct ScreenV { float x, y, z; };   // window coordinates (pixel centres at +0.5), depth in [0,1]

__device__ __forceinline__ ScreenV to_screen(const ClipV& c, int W, int H) {
    const float iw = __frcp_rn(c.w);
    ScreenV s;
    s.x = fmul(fadd(fmul(fmul(c.x, iw), 0.5f), 0.5f), (float)W);
    s.y = fmul(fadd(fmul(fmul(c.y, iw), 0.5f), 0.5f), (float)H);
    s.z = fadd(fmul(fmul(c.z, iw), 0.5f), 0.5f);
    return s;
}

// Edge function with canonical endpoint order so that the two triangles sharing an edge see exactly
// opposite values (watertight, no double hits).
__device__ __forceinline__ float edge_fn(const ScreenV& a, const ScreenV& b, float px, float py) {
    const bool swap = (a.x > b.x) || (a.x == b.x && a.y > b.y);
    const ScreenV& p = swap ? b : a;
    const ScreenV& q = swap ? a : b;
    const float e = fadd(fmul(fadd(q.x, -p.x), fadd(py, -p.y)), -fmul(fadd(q.y, -p.y), fadd(px, -p.x)));
    return swap ? -e : e;
}

// Top-left ownership of an edge a->b of a triangle normalised to positive edge_fn area (row index
// grows with y): a horizontal edge with the interior at larger y has dx > 0 ("top"), an edge with
// the interior at larger x has dy < 0 ("left").  A shared edge is traversed in opposite directions by
// its two triangles, so exactly one of them owns the pixels lying exactly on it.
__device__ __forceinline__ bool is_top_left(const ScreenV& a, const ScreenV& b) {
    const float dx = fadd(b.x, -a.x), dy = fadd(b.y, -a.y);
    return (dy == 0.f && dx > 0.f) || (dy < 0.f);
}

template <int kGroup>
__device__ void raster_triangle(ScreenV v0, ScreenV v1, ScreenV v2, int W, int H, int face, int lane,
                                int area_lo, int area_hi, u64* __restrict__ zbuf) {
    float area = edge_fn(v0, v1, v2.x, v2.y);
    if (area == 0.f || !(area == area)) return;
    if (area < 0.f) {   // make orientation positive
        ScreenV t = v1; v1 = v2; v2 = t;
        area = -area;
    }
    const float minx = fminf(v0.x, fminf(v1.x, v2.x)), maxx = fmaxf(v0.x, fmaxf(v1.x, v2.x));
    const float miny = fminf(v0.y, fminf(v1.y, v2.y)), maxy = fmaxf(v0.y, fmaxf(v1.y, v2.y));
    // pixels whose centre (i+0.5) can lie inside [min, max]
    int x0 = max(0, (int)ceilf(minx - 0.5f)), x1 = min(W - 1, (int)floorf(maxx - 0.5f));
    int y0 = max(0, (int)ceilf(miny - 0.5f)), y1 = min(H - 1, (int)floorf(maxy - 0.5f));
    if (x0 > x1 || y0 > y1) return;
    const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
    const long long barea = (long long)bw * bh;
    if (barea <= area_lo || barea > area_hi) return;
    const bool tl0 = is_top_left(v1, v2), tl1 = is_top_left(v2, v0), tl2 = is_top_left(v0, v1);
    const float inv_area = __frcp_rn(area);
    for (long long k = lane; k < barea; k += kGroup) {
        const int px = x0 + (int)(k % bw), py = y0 + (int)(k / bw);
        const float cx = fadd((float)px, 0.5f), cy = fadd((float)py, 0.5f);
        const float w0 = edge_fn(v1, v2, cx, cy);
        const float w1 = edge_fn(v2, v0, cx, cy);
        const float w2 = edge_fn(v0, v1, cx, cy);
        const bool in0 = w0 > 0.f || (w0 == 0.f && tl0);
        const bool in1 = w1 > 0.f || (w1 == 0.f && tl1);
        const bool in2 = w2 > 0.f || (w2 == 0.f && tl2);
        if (!(in0 && in1 && in2)) continue;
        const float z = fmul(fadd(fadd(fmul(w0, v0.z), fmul(w1, v1.z)), fmul(w2, v2.z)), inv_area);
        if (!(z >= 0.f && z <= 1.f)) continue;   // per-fragment near/far clip (-w <= z_clip <= w)
        const u64 key = ((u64)__float_as_uint(z) << 32) | (uint32_t)face;
        atomicMin(zbuf + (size_t)py * W + px, key);
    }
}

// Clip-space setup of one face: up to two screen triangles (near-plane clipping by Sutherland-Hodgman).
struct FaceTris {
    ScreenV t[2][3];
    int n;
};

__device__ __forceinline__ FaceTris setup_face(long long f, const float* __restrict__ verts,
                                               const int32_t* __restrict__ faces, const float* m, int W, int H) {
    FaceTris out;
    out.n = 0;
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    ClipV c[3];
    c[0] = to_clip(m, verts[3 * (size_t)i0], verts[3 * (size_t)i0 + 1], verts[3 * (size_t)i0 + 2]);
    c[1] = to_clip(m, verts[3 * (size_t)i1], verts[3 * (size_t)i1 + 1], verts[3 * (size_t)i1 + 2]);
    c[2] = to_clip(m, verts[3 * (size_t)i2], verts[3 * (size_t)i2 + 1], verts[3 * (size_t)i2 + 2]);
    // near plane of the GL clip volume: z + w >= 0 (implies w > 0 for any sane projection)
    const float d0 = fadd(c[0].z, c[0].w), d1 = fadd(c[1].z, c[1].w), d2 = fadd(c[2].z, c[2].w);
    const bool in0 = d0 >= 0.f && c[0].w > 1e-12f, in1 = d1 >= 0.f && c[1].w > 1e-12f, in2 = d2 >= 0.f && c[2].w > 1e-12f;
    const int nin = (int)in0 + (int)in1 + (int)in2;
    if (nin == 0) return out;
    if (nin == 3) {
        out.t[0][0] = to_screen(c[0], W, H); out.t[0][1] = to_screen(c[1], W, H); out.t[0][2] = to_screen(c[2], W, H);
        out.n = 1;
        return out;
    }
    // Sutherland-Hodgman against the near plane: up to 4 vertices
    ClipV poly[4];
    int np = 0;
    const float d[3] = {d0, d1, d2};
    const bool in[3] = {in0, in1, in2};
    for (int e = 0; e < 3; ++e) {
        const int a = e, b = (e + 1) % 3;
        if (in[a]) poly[np++] = c[a];
        if (in[a] != in[b]) {
            const float t = __fdiv_rn(d[a], fadd(d[a], -d[b]));
            ClipV p = lerp(c[a], c[b], t);
            if (!(p.w > 1e-12f)) p.w = 1e-12f;
            poly[np++] = p;
        }
    }
    if (np < 3) return out;
    const ScreenV s0 = to_screen(poly[0], W, H), s1 = to_screen(poly[1], W, H), s2 = to_screen(poly[2], W, H);
    out.t[0][0] = s0; out.t[0][1] = s1; out.t[0][2] = s2;
    out.n = 1;
    if (np == 4) {
        out.t[1][0] = s0; out.t[1][1] = s2; out.t[1][2] = to_screen(poly[3], W, H);
        out.n = 2;
    }
    return out;
}

// Pixel-centre bounding-box area of a screen triangle (0 if it cannot cover a pixel centre), the size-class key.
__device__ __forceinline__ long long bbox_area(const ScreenV& v0, const ScreenV& v1, const ScreenV& v2, int W, int H) {
    const float minx = fminf(v0.x, fminf(v1.x, v2.x)), maxx = fmaxf(v0.x, fmaxf(v1.x, v2.x));
    const float miny = fminf(v0.y, fminf(v1.y, v2.y)), maxy = fmaxf(v0.y, fmaxf(v1.y, v2.y));
    const int x0 = max(0, (int)ceilf(minx - 0.5f)), x1 = min(W - 1, (int)floorf(maxx - 0.5f));
    const int y0 = max(0, (int)ceilf(miny - 0.5f)), y1 = min(H - 1, (int)floorf(maxy - 0.5f));
    if (x0 > x1 || y0 > y1) return 0;
    return (long long)(x1 - x0 + 1) * (y1 - y0 + 1);
}

constexpr int kSmallArea = 32, kMidArea = 8192;

// Pass 1, one thread per face: small triangles (the overwhelming majority at ~1 M faces / 1080p) are rasterised on the
// spot; a face with a larger triangle goes onto the warp list or the CTA list and is rasterised by pass 2 over that list
// ONLY.  (Round 1 launched the warp and CTA classes over all F faces -- F x 256 threads that loaded, projected and
// exited, mesh_vis.cu r1:206-221.)  lists = [n_mid, n_big, pad, pad | mid faces (F) | big faces (F)].
__global__ void __launch_bounds__(256)
classify_faces_kernel(int F, const float* __restrict__ verts, const int32_t* __restrict__ faces,
                      const float* __restrict__ proj, int W, int H, u64* __restrict__ zbuf, int32_t* __restrict__ lists) {
    __shared__ float m[16];
    if (threadIdx.x < 16) m[threadIdx.x] = proj[threadIdx.x];
    __syncthreads();
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const FaceTris ft = setup_face(f, verts, faces, m, W, H);
    long long amax = 0;
    for (int k = 0; k < ft.n; ++k) amax = max(amax, bbox_area(ft.t[k][0], ft.t[k][1], ft.t[k][2], W, H));
    if (amax == 0) return;
    if (amax <= kSmallArea) {
        for (int k = 0; k < ft.n; ++k)
            raster_triangle<1>(ft.t[k][0], ft.t[k][1], ft.t[k][2], W, H, (int)f, 0, 0, 0x7fffffff, zbuf);
    } else if (amax <= kMidArea) {
        lists[4 + atomicAdd(lists + 0, 1)] = (int32_t)f;
    } else {
        lists[4 + F + atomicAdd(lists + 1, 1)] = (int32_t)f;
    }
}

// Pass 2: kGroup threads per LISTED face (32: a warp, 256: a CTA), striding over the list.
template <int kGroup>
__global__ void __launch_bounds__(256)
raster_listed_faces_kernel(int F, const float* __restrict__ verts, const int32_t* __restrict__ faces,
                           const float* __restrict__ proj, int W, int H, u64* __restrict__ zbuf,
                           const int32_t* __restrict__ count, const int32_t* __restrict__ list) {
    __shared__ float m[16];
    if (threadIdx.x < 16) m[threadIdx.x] = proj[threadIdx.x];
    __syncthreads();
    const int n = *count;
    const int groups_per_block = 256 / kGroup;
    const int lane = threadIdx.x % kGroup;
    for (long long i = (long long)blockIdx.x * groups_per_block + threadIdx.x / kGroup; i < n;
         i += (long long)gridDim.x * groups_per_block) {
        const long long f = list[i];
        const FaceTris ft = setup_face(f, verts, faces, m, W, H);
        for (int k = 0; k < ft.n; ++k)
            raster_triangle<kGroup>(ft.t[k][0], ft.t[k][1], ft.t[k][2], W, H, (int)f, lane, 0, 0x7fffffff, zbuf);
    }
}

// Without scratch space for the lists (d_scratch == NULL): every size class is launched over all faces.
template <int kGroup>
__global__ void __launch_bounds__(256)
raster_faces_kernel(int F, const float* __restrict__ verts, const int32_t* __restrict__ faces,
                    const float* __restrict__ proj, int W, int H, int area_lo, int area_hi,
                    u64* __restrict__ zbuf) {
    __shared__ float m[16];
    if (threadIdx.x < 16) m[threadIdx.x] = proj[threadIdx.x];
    __syncthreads();
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long f = gtid / kGroup;
    const int lane = (int)(gtid % kGroup);
    if (f >= F) return;
    const FaceTris ft = setup_face(f, verts, faces, m, W, H);
    for (int k = 0; k < ft.n; ++k)
        raster_triangle<kGroup>(ft.t[k][0], ft.t[k][1], ft.t[k][2], W, H, (int)f, lane, area_lo, area_hi, zbuf);
}

__global__ void __launch_bounds__(256)
resolve_kernel(int N, int F, const u64* __restrict__ zbuf, int32_t* __restrict__ pix_to_face,
               uint8_t* __restrict__ face_visible, int mark_last_on_bg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const u64 k = zbuf[i];
    int face = -1;
    if (k != ~0ull) face = (int)(uint32_t)k;
    pix_to_face[i] = face;
    if (face_visible != nullptr) {
        if (face >= 0) face_visible[face] = 1;
        else if (mark_last_on_bg && F > 0) face_visible[F - 1] = 1;
    }
}

__global__ void __launch_bounds__(256)
mask_from_faces_kernel(int n_points, const long long* __restrict__ cells, int F,
                       const uint8_t* __restrict__ face_visible, int n_bg, uint8_t* __restrict__ mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_points + n_bg) return;
    if (i < n_points) {
        long long c = cells[i];
        if (c < 0) c += F;   // torch negative indexing
        mask[i] = (c >= 0 && c < F) ? face_visible[c] : 0;
    } else {
        mask[i] = 1;
    }
}

}  // namespace

cudaError_t launch_mesh_visibility(int V, int F, const float* verts, const int32_t* faces, const float* proj,
                                   int W, int H, unsigned long long* zbuf, int32_t* pix_to_face,
                                   uint8_t* face_visible, int mark_last_on_bg, int32_t* scratch, cudaStream_t s) {
    (void)V;
    const size_t N = (size_t)W * H;
    cudaError_t e = cudaMemsetAsync(zbuf, 0xff, N * 8, s);
    if (e != cudaSuccess) return e;
    if (face_visible != nullptr && F > 0) {
        e = cudaMemsetAsync(face_visible, 0, (size_t)F, s);
        if (e != cudaSuccess) return e;
    }
    if (F > 0 && scratch != nullptr) {
        e = cudaMemsetAsync(scratch, 0, 16, s);
        if (e != cudaSuccess) return e;
        classify_faces_kernel<<<(unsigned)(((long long)F + 255) / 256), 256, 0, s>>>(F, verts, faces, proj, W, H, zbuf,
                                                                                     scratch);
        // the lists live on the device: a fixed, SM-sized grid strides over whatever they hold (usually a few faces)
        raster_listed_faces_kernel<32><<<148 * 2, 256, 0, s>>>(F, verts, faces, proj, W, H, zbuf, scratch + 0,
                                                               scratch + 4);
        raster_listed_faces_kernel<256><<<148, 256, 0, s>>>(F, verts, faces, proj, W, H, zbuf, scratch + 1,
                                                            scratch + 4 + F);
        count_launch(3);
    } else if (F > 0) {
        {
            const long long threads = (long long)F;
            raster_faces_kernel<1><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(
                F, verts, faces, proj, W, H, 0, kSmallArea, zbuf);
        }
        {
            const long long threads = (long long)F * 32;
            raster_faces_kernel<32><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(
                F, verts, faces, proj, W, H, kSmallArea, kMidArea, zbuf);
        }
        {
            const long long threads = (long long)F * 256;
            raster_faces_kernel<256><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(
                F, verts, faces, proj, W, H, kMidArea, 0x7fffffff, zbuf);
        }
        count_launch(3);
    }
    if (N > 0) count_launch();
    if (N > 0)
        resolve_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>((int)N, F, zbuf, pix_to_face, face_visible,
                                                                   mark_last_on_bg);
    return cudaGetLastError();
}

cudaError_t launch_mask_from_faces(int n_points, const long long* cells, int F, const uint8_t* face_visible,
                                   int n_bg, uint8_t* mask, cudaStream_t s) {
    const int n = n_points + n_bg;
    if (n > 0) { mask_from_faces_kernel<<<(n + 255) / 256, 256, 0, s>>>(n_points, cells, F, face_visible, n_bg, mask); count_launch(); }
    return cudaGetLastError();
}

}  // namespace fb200
