// api.cu -- the extern "C" boundary declared in include/frosting_b200.h.
//
// Mirrors, for this path, what DGR/rasterize_points.cu + CudaRasterizer::Rasterizer do for the
// reference: argument checks, workspace carving, kernel sequencing.  Differences by design:
// raw device pointers instead of torch tensors, an explicit stream, no allocation, no host
// synchronisation unless `debug` is set.
#include <atomic>
#include <mutex>
#include <cstdio>
#include <cstring>

#include "common.cuh"

using namespace fb200;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}

int check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return FB200_OK;
    snprintf(g_err, sizeof(g_err), "CUDA error in %s: %s", what, cudaGetErrorString(e));
    return FB200_ECUDA;
}

// debug mode: synchronise after every stage like CHECK_CUDA (auxiliary.h:166-173)
int stage(cudaError_t e, const char* what, bool debug, cudaStream_t s) {
    int rc = check(e, what);
    if (rc != FB200_OK) return rc;
    if (debug) return check(cudaStreamSynchronize(s), what);
    return FB200_OK;
}

int check_frosting(const fb200_frosting_params* fp) {
    if (!fp) return fail(FB200_EINVAL, "frosting_attributes: null params%s");
    if (fp->P < 0 || fp->sh_rest < 0 || fp->n_verts < 0 || fp->n_faces < 0)
        return fail(FB200_EINVAL, "frosting_attributes: bad extents%s");
    if (fp->P > 0 && (!fp->d_bary_logits || !fp->d_cells || !fp->d_faces || !fp->d_inner_verts || !fp->d_outer_verts ||
                      !fp->d_opacity_logits || !fp->d_log_scales || !fp->d_quats || !fp->d_sh_dc ||
                      (fp->sh_rest > 0 && !fp->d_sh_rest)))
        return fail(FB200_EINVAL, "frosting_attributes: missing parameter pointer%s");
    if (fp->P > 0 && ((reinterpret_cast<uintptr_t>(fp->d_bary_logits) & 7) || (reinterpret_cast<uintptr_t>(fp->d_quats) & 15)))
        return fail(FB200_EINVAL, "frosting_attributes: bary logits must be 8-byte, quats 16-byte aligned%s");
    return FB200_OK;
}

// frosting mode (fb200_inputs.frosting): the rasterizer's own inputs are derived from the parameter block
int validate_frosting(const fb200_params* prm, const fb200_inputs* in) {
    const fb200_frosting_params* fp = in->frosting;
    int rc = check_frosting(fp);
    if (rc != FB200_OK) return rc;
    if (in->d_means3D || in->d_shs || in->d_colors_precomp || in->d_opacities || in->d_scales || in->d_rotations ||
        in->d_cov3D_precomp || in->d_visibility || in->d_point_cells || in->d_face_visible)
        return fail(FB200_EINVAL, "frosting mode: the attribute / mask pointers of fb200_inputs must be NULL%s");
    if (fp->P != prm->P || prm->sh_coeffs != fp->sh_rest + 1 || fp->sh_rest > 15)
        return fail(FB200_EINVAL, "frosting mode: P / sh_coeffs do not match the parameter block (sh_rest <= 15)%s");
    if (prm->sh_degree < 0 || prm->sh_degree > 3 || (prm->sh_degree + 1) * (prm->sh_degree + 1) > prm->sh_coeffs)
        return fail(FB200_EINVAL, "sh_degree / sh_coeffs inconsistent%s");
    if ((reinterpret_cast<uintptr_t>(fp->d_quats) & 15) || (reinterpret_cast<uintptr_t>(fp->d_bary_logits) & 7))
        return fail(FB200_EINVAL, "frosting mode: quats must be 16-byte aligned, bary logits 8-byte aligned%s");
    if (prm->extra) return fail(FB200_EINVAL, "frosting mode: extra feature channels are not supported%s");
    return FB200_OK;
}

// the fb200_inputs the kernels see in frosting mode: culling inputs taken from the parameter block
fb200_inputs frosting_inputs(const fb200_inputs* in) {
    fb200_inputs r = *in;
    const fb200_frosting_params* fp = in->frosting;
    r.d_visibility = fp->d_mask;
    if (fp->d_face_visible) { r.d_point_cells = fp->d_cells; r.d_face_visible = fp->d_face_visible; r.n_cell_points = fp->P; }
    r.frosting = nullptr;     // a host pointer: never dereferenced on the device
    return r;
}

int validate(const fb200_params* prm, const fb200_inputs* in, const fb200_workspace* ws) {
    if (!prm || !in || !ws) return fail(FB200_EINVAL, "null argument struct%s");
    if (prm->P < 0 || prm->image_width <= 0 || prm->image_height <= 0)
        return fail(FB200_EINVAL, "bad extents (P >= 0, image dims > 0 required)%s");
    if (prm->image_width > 65535 * FB200_TILE || prm->image_height > 65535 * FB200_TILE)
        return fail(FB200_EINVAL, "image too large for 16-bit tile coordinates%s");
    if (in->frosting) {
        int rc = validate_frosting(prm, in);
        if (rc != FB200_OK) return rc;
    } else if (prm->P > 0) {
        if (!in->d_means3D || !in->d_opacities) return fail(FB200_EINVAL, "means3D / opacities missing%s");
        const bool has_sh = in->d_shs != nullptr, has_col = in->d_colors_precomp != nullptr;
        if (has_sh == has_col)
            return fail(FB200_EINVAL, "Please provide excatly one of either SHs or precomputed colors!%s");
        const bool has_sr = in->d_scales != nullptr && in->d_rotations != nullptr;
        const bool any_sr = in->d_scales != nullptr || in->d_rotations != nullptr;
        const bool has_cov = in->d_cov3D_precomp != nullptr;
        if ((!has_sr && !has_cov) || (any_sr && has_cov))
            return fail(FB200_EINVAL,
                        "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!%s");
        if (has_sh && (prm->sh_coeffs <= 0 || prm->sh_degree < 0 || prm->sh_degree > 3 ||
                       (prm->sh_degree + 1) * (prm->sh_degree + 1) > prm->sh_coeffs))
            return fail(FB200_EINVAL, "sh_degree / sh_coeffs inconsistent%s");
        // rows read with 128-bit loads (header: Alignment)
        if ((in->d_rotations && (reinterpret_cast<uintptr_t>(in->d_rotations) & 15)) ||
            (has_sh && (prm->sh_coeffs * 3) % 4 == 0 && (reinterpret_cast<uintptr_t>(in->d_shs) & 15)))
            return fail(FB200_EINVAL, "rotations / SH rows must be 16-byte aligned%s");
        if ((in->d_point_cells == nullptr) != (in->d_face_visible == nullptr))
            return fail(FB200_EINVAL, "give both d_point_cells and d_face_visible or neither%s");
        if (in->d_point_cells && (in->n_cell_points < 0 || in->n_cell_points > prm->P))
            return fail(FB200_EINVAL, "n_cell_points must be within [0, P]%s");
    }
    if (!in->d_background || !in->d_viewmatrix || !in->d_projmatrix || !in->d_campos)
        return fail(FB200_EINVAL, "camera tensors missing%s");
    if (!ws->d_geom || !ws->d_image || !ws->d_status) return fail(FB200_EINVAL, "workspace pointers missing%s");
    if (ws->binning_capacity > 0 && !ws->d_binning) return fail(FB200_EINVAL, "binning buffer missing%s");
    if (ws->binning_capacity < 0 || ws->binning_capacity > 0x7fffffffLL)
        return fail(FB200_EINVAL, "binning capacity out of range%s");
    const GeomLayout gl((size_t)prm->P);
    const ImageLayout il(prm->image_width, prm->image_height);
    const BinLayout bl((size_t)ws->binning_capacity);
    if (ws->geom_bytes < gl.total) return fail(FB200_ENOSPC, "geometry workspace too small%s");
    if (ws->image_bytes < il.total) return fail(FB200_ENOSPC, "image workspace too small%s");
    if (ws->binning_capacity > 0 && ws->binning_bytes < bl.total)
        return fail(FB200_ENOSPC, "binning workspace too small%s");
    return FB200_OK;
}

// row f4: copy the caller's extra-channel block into kernel arguments
int setup_extra(const fb200_params* prm, bool forward, bool backward, ExtraArgs& ex) {
    memset(&ex, 0, sizeof(ex));
    const fb200_extra* e = prm->extra;
    if (!e || prm->P == 0) return FB200_OK;
    if (e->channels < 1 || e->channels > FB200_CHANNELS)
        return fail(FB200_EINVAL, "extra feature channels must be 1..3%s");
    if (!e->d_features || !e->d_background) return fail(FB200_EINVAL, "extra features / background missing%s");
    if (forward && !e->d_out) return fail(FB200_EINVAL, "extra output image missing%s");
    if (backward && (!e->d_dL_dout || !e->d_dL_dfeatures)) return fail(FB200_EINVAL, "extra gradient pointers missing%s");
    ex.ch = e->channels; ex.feat = e->d_features; ex.bg = e->d_background; ex.out = e->d_out;
    ex.dL_dout = e->d_dL_dout; ex.dL_dfeat = e->d_dL_dfeatures;
    return FB200_OK;
}

// ---- measurement hooks -------------------------------------------------------------------------------
std::atomic<long long> g_launches{0};
std::atomic<int> g_profile{0};
// Process-wide (not thread-local): torch's autograd engine calls fb200_backward from its own thread.
struct StageEvents {
    cudaEvent_t ev[FB200_NUM_STAGES][2];
    bool made = false;
    bool recorded[FB200_NUM_STAGES] = {};
};
StageEvents g_ev;
std::mutex g_ev_mu;

struct StageTimer {
    int stage;
    cudaStream_t s;
    bool on;
    StageTimer(int st, cudaStream_t stream) : stage(st), s(stream), on(g_profile.load() != 0) {
        if (!on) return;
        std::lock_guard<std::mutex> lk(g_ev_mu);
        if (!g_ev.made) {
            for (int i = 0; i < FB200_NUM_STAGES; ++i) { cudaEventCreate(&g_ev.ev[i][0]); cudaEventCreate(&g_ev.ev[i][1]); }
            g_ev.made = true;
        }
        cudaEventRecord(g_ev.ev[stage][0], s);
    }
    ~StageTimer() {
        if (!on) return;
        std::lock_guard<std::mutex> lk(g_ev_mu);
        cudaEventRecord(g_ev.ev[stage][1], s);
        g_ev.recorded[stage] = true;
    }
};

inline char* align128(void* p) {
    return reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 127) & ~uintptr_t(127));
}

}  // namespace

namespace fb200 {
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace fb200

extern "C" {

int fb200_profile_enable(int32_t enable) { g_profile.store(enable ? 1 : 0); return FB200_OK; }

int fb200_profile_read(float* ms_out) {
    if (!ms_out) return fail(FB200_EINVAL, "profile_read: null output%s");
    std::lock_guard<std::mutex> lk(g_ev_mu);
    for (int i = 0; i < FB200_NUM_STAGES; ++i) {
        ms_out[i] = -1.0f;
        if (g_ev.made && g_ev.recorded[i]) {
            cudaError_t e = cudaEventSynchronize(g_ev.ev[i][1]);
            if (e != cudaSuccess) return check(e, "profile_read");
            e = cudaEventElapsedTime(&ms_out[i], g_ev.ev[i][0], g_ev.ev[i][1]);
            if (e != cudaSuccess) return check(e, "profile_read");
        }
    }
    return FB200_OK;
}

int64_t fb200_kernel_launches(void) { return g_launches.load(); }

int fb200_abi_version(void) { return FB200_ABI_VERSION; }

int fb200_abi_struct_sizes(size_t* out) {
    if (!out) return fail(FB200_EINVAL, "abi_struct_sizes: null output%s");
    const size_t sizes[FB200_ABI_STRUCTS] = {sizeof(fb200_params), sizeof(fb200_inputs), sizeof(fb200_workspace),
                                             sizeof(fb200_grads), sizeof(fb200_extra), sizeof(fb200_frosting_params),
                                             sizeof(fb200_frosting_grads), sizeof(fb200_adam_args), sizeof(fb200_layout)};
    for (int i = 0; i < FB200_ABI_STRUCTS; ++i) out[i] = sizes[i];
    return FB200_OK;
}
const char* fb200_last_error(void) { return g_err; }

size_t fb200_geom_bytes(int32_t P) { return GeomLayout((size_t)(P < 0 ? 0 : P)).total; }
size_t fb200_image_bytes(int32_t W, int32_t H) { return ImageLayout(W < 1 ? 1 : W, H < 1 ? 1 : H).total; }
size_t fb200_binning_bytes(int64_t capacity) { return BinLayout((size_t)(capacity < 0 ? 0 : capacity)).total; }
size_t fb200_rec_stream_bytes(int64_t capacity) { return (size_t)(capacity < 0 ? 0 : capacity) * sizeof(SplatRec) + 256; }

int fb200_get_layout(int32_t P, int32_t W, int32_t H, int64_t capacity, fb200_layout* out) {
    if (!out || P < 0 || W <= 0 || H <= 0 || capacity < 0) return fail(FB200_EINVAL, "bad layout query%s");
    const GeomLayout gl((size_t)P);
    const ImageLayout il(W, H);
    const BinLayout bl((size_t)capacity);
    out->geom_rec = gl.rec; out->geom_depth = gl.depth; out->geom_rect = gl.rect; out->geom_clamped = gl.clamped;
    out->img_final_T = il.final_T; out->img_n_contrib = il.n_contrib; out->img_ranges = il.ranges;
    out->img_tile_count = il.tile_count;
    out->bin_point_list = bl.point_list; out->bin_keys = bl.keys;
    return FB200_OK;
}

static int setup_fwd(const fb200_params* prm, const fb200_inputs* in, const fb200_workspace* ws, bool need_binning,
                     float* d_out_color, int32_t* d_radii, FwdArgs& a) {
    int rc = validate(prm, in, ws);
    if (rc != FB200_OK) return rc;
    if (prm->P > 0 && !d_radii) return fail(FB200_EINVAL, "output pointers missing%s");
    if (need_binning && !d_out_color) return fail(FB200_EINVAL, "output pointers missing%s");
    const GeomLayout gl((size_t)prm->P);
    const ImageLayout il(prm->image_width, prm->image_height);
    const BinLayout bl((size_t)ws->binning_capacity);
    char* g = align128(ws->d_geom);
    char* im = align128(ws->d_image);
    char* bn = ws->d_binning ? align128(ws->d_binning) : nullptr;
    memset(&a, 0, sizeof(a));
    a.prm = *prm;
    a.in = *in;
    if (in->frosting) { a.frosting = 1; a.fr = *in->frosting; a.in = frosting_inputs(in); }
    // focal lengths exactly as rasterizer_impl.cu:222-223
    a.focal_y = prm->image_height / (2.0f * prm->tanfovy);
    a.focal_x = prm->image_width / (2.0f * prm->tanfovx);
    a.tiles_x = il.tiles_x; a.tiles_y = il.tiles_y;
    a.rec = reinterpret_cast<SplatRec*>(g + gl.rec);
    a.depth = reinterpret_cast<float*>(g + gl.depth);
    a.rect = reinterpret_cast<uint2*>(g + gl.rect);
    a.clamped = reinterpret_cast<uint8_t*>(g + gl.clamped);
    a.vis_list = reinterpret_cast<uint32_t*>(g + gl.vis_list);
    a.final_T = reinterpret_cast<float*>(im + il.final_T);
    a.n_contrib = reinterpret_cast<uint32_t*>(im + il.n_contrib);
    a.last_entry = reinterpret_cast<uint32_t*>(im + il.last_entry);
    a.ranges = reinterpret_cast<uint2*>(im + il.ranges);
    a.tile_count = reinterpret_cast<uint32_t*>(im + il.tile_count);
    a.cursor = reinterpret_cast<uint32_t*>(im + il.cursor);
    a.list_tiny = reinterpret_cast<uint32_t*>(im + il.list_tiny);
    a.list_small = reinterpret_cast<uint32_t*>(im + il.list_small);
    a.list_large = reinterpret_cast<uint32_t*>(im + il.list_large);
    a.list_huge = reinterpret_cast<uint32_t*>(im + il.list_huge);
    a.tile_order = reinterpret_cast<uint32_t*>(im + il.tile_order);
    a.counters = reinterpret_cast<uint32_t*>(im + il.counters);
    a.point_list = bn ? reinterpret_cast<uint32_t*>(bn + bl.point_list) : nullptr;
    a.keys = bn ? reinterpret_cast<unsigned long long*>(bn + bl.keys) : nullptr;
    a.keys_scratch = bn ? reinterpret_cast<unsigned long long*>(bn + bl.keys_scratch) : nullptr;
    a.sub_hits = bn ? reinterpret_cast<uint32_t*>(bn + bl.sub_hits) : nullptr;
    a.capacity = ws->binning_capacity;
    a.rec_stream = nullptr;
    if ((prm->debug & 8) && ws->d_rec_stream) {
        if (ws->rec_stream_bytes < fb200_rec_stream_bytes(ws->binning_capacity))
            return fail(FB200_ENOSPC, "record stream workspace too small%s");
        a.rec_stream = reinterpret_cast<SplatRec*>(align128(ws->d_rec_stream));
    }
    a.status = ws->d_status;
    a.out_color = d_out_color;
    a.radii = d_radii;
    return FB200_OK;
}

int fb200_forward_geometry(const fb200_params* prm, const fb200_inputs* in, const fb200_workspace* ws,
                           int32_t* d_radii, void* stream) {
    FwdArgs a;
    int rc = setup_fwd(prm, in, ws, false, nullptr, d_radii, a);
    if (rc != FB200_OK) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool debug = (prm->debug & 1) != 0;
    { StageTimer t(FB200_STAGE_PREPROCESS, s);
      if ((rc = stage(launch_preprocess_fwd(a, s), "preprocess", debug, s)) != FB200_OK) return rc;
      if ((rc = stage(launch_tile_scan(a, s), "tile scan", debug, s)) != FB200_OK) return rc; }
    if (ws->acc_zeroed_by_forward && prm->P > 0) {
        // backward's accumulators: cleared here, in the shadow of the host's wait for the instance count
        const GeomLayout gl((size_t)prm->P);
        float* acc = reinterpret_cast<float*>(align128(ws->d_geom) + gl.acc);
        if ((rc = stage(cudaMemsetAsync(acc, 0, sizeof(float) * 12 * (size_t)prm->P, s), "accumulator clear", debug, s)) != FB200_OK)
            return rc;
    }
    return FB200_OK;
}

int fb200_forward_raster(const fb200_params* prm, const fb200_inputs* in, const fb200_workspace* ws,
                         float* d_out_color, int32_t* d_radii, void* stream) {
    FwdArgs a;
    int rc = setup_fwd(prm, in, ws, true, d_out_color, d_radii, a);
    if (rc != FB200_OK) return rc;
    if ((rc = setup_extra(prm, true, false, a.ex)) != FB200_OK) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool debug = (prm->debug & 1) != 0;
    { StageTimer t(FB200_STAGE_BINNING, s);
      if ((rc = stage(launch_binning(a, s, ws->h_status), "binning", debug, s)) != FB200_OK) return rc; }
    { StageTimer t(FB200_STAGE_RENDER_FWD, s);
      if ((rc = stage(launch_render_fwd(a, s), "render", debug, s)) != FB200_OK) return rc; }
    return FB200_OK;
}

int fb200_forward(const fb200_params* prm, const fb200_inputs* in, const fb200_workspace* ws,
                  float* d_out_color, int32_t* d_radii, void* stream) {
    int rc = fb200_forward_geometry(prm, in, ws, d_radii, stream);
    if (rc != FB200_OK) return rc;
    if (!ws) return fail(FB200_EINVAL, "null argument struct%s");
    fb200_workspace one_phase = *ws;
    one_phase.h_status = nullptr;      // nobody has synchronised: the host cannot know this frame's status words
    return fb200_forward_raster(prm, in, &one_phase, d_out_color, d_radii, stream);
}

int fb200_backward(const fb200_params* prm, const fb200_inputs* in, const fb200_workspace* ws,
                   const int32_t* d_radii, const float* d_dL_dout_color, const fb200_grads* grads, void* stream) {
    int rc = validate(prm, in, ws);
    if (rc != FB200_OK) return rc;
    if (!grads || !d_dL_dout_color || (prm->P > 0 && !d_radii)) return fail(FB200_EINVAL, "backward pointers missing%s");
    // outputs a caller cannot use may be NULL and are then not written: dL/dcolors on the SH path (an intermediate
    // there), dL/dcov3D on the scale/rotation path, dL/dscales + dL/drotations on the precomputed-covariance path
    const fb200_frosting_grads* fg = grads->frosting;
    if (in->frosting) {
        if (!fg) return fail(FB200_EINVAL, "frosting mode: fb200_grads.frosting missing%s");
        if (prm->P > 0 && (!fg->d_bary_logits || !fg->d_opacity_logits || !fg->d_log_scales || !fg->d_quats || !fg->d_sh_dc ||
                           (in->frosting->sh_rest > 0 && !fg->d_sh_rest) ||
                           ((fg->d_inner_verts == nullptr) != (fg->d_outer_verts == nullptr))))
            return fail(FB200_EINVAL, "frosting mode: parameter-gradient pointers missing%s");
        if ((reinterpret_cast<uintptr_t>(fg->d_quats) & 15) || (reinterpret_cast<uintptr_t>(fg->d_bary_logits) & 7))
            return fail(FB200_EINVAL, "frosting mode: dL/dquats must be 16-byte aligned, dL/dbary 8-byte aligned%s");
    } else if (fg) {
        return fail(FB200_EINVAL, "fb200_grads.frosting given without fb200_inputs.frosting%s");
    } else if (prm->P > 0 && (!grads->d_dL_dmeans2D || !grads->d_dL_dopacity || !grads->d_dL_dmeans3D ||
                       (in->d_colors_precomp && !grads->d_dL_dcolors) || (in->d_cov3D_precomp && !grads->d_dL_dcov3D) ||
                       (!in->d_cov3D_precomp && (!grads->d_dL_dscales || !grads->d_dL_drotations)) ||
                       (in->d_shs && prm->sh_coeffs > 0 && !grads->d_dL_dsh)))
        return fail(FB200_EINVAL, "gradient output pointers missing%s");
    if ((grads->d_dL_drotations && (reinterpret_cast<uintptr_t>(grads->d_dL_drotations) & 15)) ||
        (grads->d_dL_dsh && (prm->sh_coeffs * 3) % 4 == 0 && (reinterpret_cast<uintptr_t>(grads->d_dL_dsh) & 15)))
        return fail(FB200_EINVAL, "dL/drotations / dL/dsh rows must be 16-byte aligned%s");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool debug = (prm->debug & 1) != 0;
    if (prm->P == 0) return FB200_OK;

    const GeomLayout gl((size_t)prm->P);
    const ImageLayout il(prm->image_width, prm->image_height);
    const BinLayout bl((size_t)ws->binning_capacity);
    char* g = align128(ws->d_geom);
    char* im = align128(ws->d_image);
    char* bn = ws->d_binning ? align128(ws->d_binning) : nullptr;

    BwdArgs a;
    memset(&a, 0, sizeof(a));
    a.prm = *prm;
    a.in = *in;
    if (in->frosting) { a.frosting = 1; a.fr = *in->frosting; a.fg = *fg; a.in = frosting_inputs(in); }
    a.focal_y = prm->image_height / (2.0f * prm->tanfovy);
    a.focal_x = prm->image_width / (2.0f * prm->tanfovx);
    a.tiles_x = il.tiles_x; a.tiles_y = il.tiles_y;
    a.rec = reinterpret_cast<const SplatRec*>(g + gl.rec);
    a.clamped = reinterpret_cast<const uint8_t*>(g + gl.clamped);
    a.vis_list = reinterpret_cast<const uint32_t*>(g + gl.vis_list);
    a.acc = reinterpret_cast<float*>(g + gl.acc);
    a.final_T = reinterpret_cast<const float*>(im + il.final_T);
    a.n_contrib = reinterpret_cast<const uint32_t*>(im + il.n_contrib);
    a.last_entry = reinterpret_cast<const uint32_t*>(im + il.last_entry);
    a.tile_order = reinterpret_cast<const uint32_t*>(im + il.tile_order);
    a.ranges = reinterpret_cast<const uint2*>(im + il.ranges);
    a.point_list = bn ? reinterpret_cast<const uint32_t*>(bn + bl.point_list) : nullptr;
    a.sub_hits = bn ? reinterpret_cast<const uint32_t*>(bn + bl.sub_hits) : nullptr;
    a.status = ws->d_status;
    a.radii = d_radii;
    a.dL_dpix = d_dL_dout_color;
    a.g = *grads;
    if ((rc = setup_extra(prm, false, true, a.ex)) != FB200_OK) return rc;

    if (!ws->acc_zeroed_by_forward &&
        (rc = stage(launch_render_bwd_clear(a, s), "render backward (clear)", debug, s)) != FB200_OK) return rc;
    // the zero rows of the dense-gradient contract are written on a side stream while the blend backward runs
    // (opt-in, debug bit 5: measured SLOWER on C3 / C5 -- 0.42 + 0.135 vs 0.344 + 0.163 ms -- the fill kernel's CTAs take
    // issue slots from the blend backward, which is issue-bound; kept for frames whose blend backward is short)
    SideStream* side = (!grads->sparse_rows && !in->frosting && !debug && (prm->debug & 32) && prm->P >= 4096) ? side_stream() : nullptr;
    std::unique_lock<std::mutex> side_use;
    if (side) {
        side_use = std::unique_lock<std::mutex>(side->use);
        if ((rc = check(cudaEventRecord(side->fork, s), "fork")) != FB200_OK) return rc;
        if ((rc = check(cudaStreamWaitEvent(side->stream, side->fork, 0), "fork")) != FB200_OK) return rc;
        if ((rc = check(launch_zero_rows(a, side->stream), "zero rows")) != FB200_OK) return rc;
        if ((rc = check(cudaEventRecord(side->join, side->stream), "join")) != FB200_OK) return rc;
        a.zeroed_elsewhere = 1;
    }
    { StageTimer t(FB200_STAGE_RENDER_BWD, s);
      if ((rc = stage(launch_render_bwd(a, s), "render backward", debug, s)) != FB200_OK) return rc; }
    if (side && (rc = check(cudaStreamWaitEvent(s, side->join, 0), "join")) != FB200_OK) return rc;
    { StageTimer t(FB200_STAGE_GEOM_BWD, s);
      if ((rc = stage(launch_geom_bwd(a, s), "geometry backward", debug, s)) != FB200_OK) return rc; }
    if ((rc = stage(launch_extra_grad(a, s), "extra feature gradients", debug, s)) != FB200_OK) return rc;
    return FB200_OK;
}

int fb200_mark_visible(int32_t P, const float* d_means3D, const float* d_viewmatrix, const float* d_projmatrix,
                       uint8_t* d_present, void* stream) {
    (void)d_projmatrix;   // the reference's x/y frustum test is commented out (auxiliary.h:154)
    if (P < 0 || (P > 0 && (!d_means3D || !d_viewmatrix || !d_present)))
        return fail(FB200_EINVAL, "mark_visible: bad arguments%s");
    return check(launch_mark_visible(P, d_means3D, d_viewmatrix, d_present, static_cast<cudaStream_t>(stream)),
                 "mark_visible");
}

int fb200_mesh_visibility(int32_t V, int32_t F, const float* d_verts, const int32_t* d_faces,
                          const float* d_full_proj, int32_t W, int32_t H, uint64_t* d_zbuf,
                          int32_t* d_pix_to_face, uint8_t* d_face_visible, int32_t mark_last_on_bg,
                          int32_t* d_scratch, void* stream) {
    if (V < 0 || F < 0 || W <= 0 || H <= 0 || !d_full_proj || !d_zbuf || !d_pix_to_face ||
        (F > 0 && (!d_verts || !d_faces)))
        return fail(FB200_EINVAL, "mesh_visibility: bad arguments%s");
    return check(launch_mesh_visibility(V, F, d_verts, d_faces, d_full_proj, W, H,
                                        reinterpret_cast<unsigned long long*>(d_zbuf), d_pix_to_face,
                                        d_face_visible, mark_last_on_bg, d_scratch, static_cast<cudaStream_t>(stream)),
                 "mesh_visibility");
}

int fb200_gaussian_mask_from_faces(int32_t n_points, const int64_t* d_point_cell_indices, int32_t F,
                                   const uint8_t* d_face_visible, int32_t n_background, uint8_t* d_mask,
                                   void* stream) {
    if (n_points < 0 || n_background < 0 || F < 0 || (n_points > 0 && (!d_point_cell_indices || !d_face_visible)) ||
        (n_points + n_background > 0 && !d_mask))
        return fail(FB200_EINVAL, "gaussian_mask_from_faces: bad arguments%s");
    return check(launch_mask_from_faces(n_points, reinterpret_cast<const long long*>(d_point_cell_indices), F,
                                        d_face_visible, n_background, d_mask, static_cast<cudaStream_t>(stream)),
                 "gaussian_mask_from_faces");
}

int fb200_frosting_attributes(const fb200_frosting_params* fp, float* d_means3D, float* d_opacities, float* d_scales,
                              float* d_rotations, float* d_shs, void* stream) {
    int rc = check_frosting(fp);
    if (rc != FB200_OK) return rc;
    if (fp->P > 0 && (!d_means3D || !d_opacities || !d_scales || !d_rotations || !d_shs))
        return fail(FB200_EINVAL, "frosting_attributes: missing output pointer%s");
    return check(launch_frosting_attr_fwd(*fp, d_means3D, d_opacities, d_scales, d_rotations, d_shs,
                                          static_cast<cudaStream_t>(stream)), "frosting_attributes");
}

int fb200_frosting_attributes_backward(const fb200_frosting_params* fp, const float* d_g_means3D,
                                       const float* d_g_opacities, const float* d_g_scales,
                                       const float* d_g_rotations, const float* d_g_shs,
                                       const fb200_frosting_grads* grads, void* stream) {
    int rc = check_frosting(fp);
    if (rc != FB200_OK) return rc;
    if (!grads) return fail(FB200_EINVAL, "frosting_attributes_backward: null grads%s");
    if (fp->P > 0 && (!d_g_means3D || !d_g_opacities || !d_g_scales || !d_g_rotations || !d_g_shs ||
                      !grads->d_bary_logits || !grads->d_opacity_logits || !grads->d_log_scales || !grads->d_quats ||
                      !grads->d_sh_dc || (fp->sh_rest > 0 && !grads->d_sh_rest) ||
                      ((grads->d_inner_verts == nullptr) != (grads->d_outer_verts == nullptr))))
        return fail(FB200_EINVAL, "frosting_attributes_backward: missing pointer%s");
    return check(launch_frosting_attr_bwd(*fp, d_g_means3D, d_g_opacities, d_g_scales, d_g_rotations, d_g_shs, *grads,
                                          static_cast<cudaStream_t>(stream)), "frosting_attributes_backward");
}

size_t fb200_loss_partials(int32_t C, int32_t H, int32_t W) {
    return (C > 0 && H > 0 && W > 0) ? l1_dssim_num_partials(C, H, W) : 0;
}

int fb200_l1_dssim_forward(const float* d_pred, const float* d_gt, int32_t C, int32_t H, int32_t W, float lambda,
                           float* d_maps, float* d_partials, float* d_loss, void* stream) {
    if (C <= 0 || H <= 0 || W <= 0 || !d_pred || !d_gt || !d_maps || !d_partials || !d_loss)
        return fail(FB200_EINVAL, "l1_dssim_forward: bad arguments%s");
    return check(launch_l1_dssim_fwd(d_pred, d_gt, C, H, W, lambda, d_maps, d_partials, d_loss,
                                     static_cast<cudaStream_t>(stream)), "l1_dssim_forward");
}

int fb200_l1_dssim_backward(const float* d_pred, const float* d_gt, const float* d_maps, int32_t C, int32_t H,
                            int32_t W, float lambda, const float* d_dL_dloss, float* d_dpred, void* stream) {
    if (C <= 0 || H <= 0 || W <= 0 || !d_pred || !d_gt || !d_maps || !d_dL_dloss || !d_dpred)
        return fail(FB200_EINVAL, "l1_dssim_backward: bad arguments%s");
    return check(launch_l1_dssim_bwd(d_pred, d_gt, d_maps, C, H, W, lambda, d_dL_dloss, d_dpred,
                                     static_cast<cudaStream_t>(stream)), "l1_dssim_backward");
}

int fb200_adam_step(const fb200_adam_args* a, void* stream) {
    if (!a) return fail(FB200_EINVAL, "adam_step: null args%s");
    if (a->world < 1 || a->world > FB200_MAX_PEERS || a->rank < 0 || a->rank >= a->world)
        return fail(FB200_EINVAL, "adam_step: world must be 1..8 and rank inside it%s");
    if (a->n_groups < 1 || a->n_groups > FB200_ADAM_MAX_GROUPS)
        return fail(FB200_EINVAL, "adam_step: 1..16 parameter groups%s");
    for (int g = 0; g <= a->n_groups; ++g)
        if ((a->group_start[g] & 3) || (g > 0 && a->group_start[g] < a->group_start[g - 1]))
            return fail(FB200_EINVAL, "adam_step: group starts must be ascending multiples of 4 elements%s");
    if ((a->shard_lo & 3) || (a->shard_hi & 3) || a->shard_lo < a->group_start[0] || a->shard_hi < a->shard_lo ||
        a->shard_hi > a->group_start[a->n_groups])
        return fail(FB200_EINVAL, "adam_step: shard must be a 4-aligned range inside the slab%s");
    if (a->shard_hi > a->shard_lo) {
        if (!a->d_exp_avg || !a->d_exp_avg_sq) return fail(FB200_EINVAL, "adam_step: missing moment buffers%s");
        for (int p = 0; p < a->world; ++p)
            if (!a->peer_params[p] || !a->peer_grads[p]) return fail(FB200_EINVAL, "adam_step: missing slab pointer%s");
    }
    if ((a->mc_grads == nullptr) != (a->mc_params == nullptr))
        return fail(FB200_EINVAL, "adam_step: give both multicast mappings or neither%s");
    {
        int given = 0;
        for (int p = 0; p < a->world; ++p) given += a->peer_row_radii[p] != nullptr;
        if (given != 0 && given != a->world) return fail(FB200_EINVAL, "adam_step: give every rank's row radii or none%s");
        if (given) {
            if (a->row_count < 1) return fail(FB200_EINVAL, "adam_step: row_count must be positive with row radii%s");
            for (int g = 0; g < a->n_groups; ++g) {
                const int64_t n = a->group_start[g + 1] - a->group_start[g];
                if (a->row_width[g] < 0 || (int64_t)a->row_width[g] * a->row_count > n || n > 0xffffffffLL)
                    return fail(FB200_EINVAL, "adam_step: row_width * row_count exceeds the group%s");
            }
        }
    }
    if (!(a->bias_correction1 > 0.f) || !(a->bias_correction2_sqrt > 0.f))
        return fail(FB200_EINVAL, "adam_step: bias corrections must be positive (step >= 1)%s");
    return check(launch_adam_shard(*a, static_cast<cudaStream_t>(stream)), "adam_step");
}

int fb200_peer_alloc(size_t bytes, void** d_ptr) {
    if (!d_ptr || bytes == 0) return fail(FB200_EINVAL, "peer_alloc: bad arguments%s");
    void* p = nullptr;
    int rc = check(cudaMalloc(&p, bytes), "peer_alloc");
    if (rc != FB200_OK) return rc;
    rc = check(cudaMemset(p, 0, bytes), "peer_alloc (zero-fill)");
    if (rc != FB200_OK) { cudaFree(p); return rc; }
    *d_ptr = p;
    return FB200_OK;
}

int fb200_peer_free(void* d_ptr) { return d_ptr ? check(cudaFree(d_ptr), "peer_free") : FB200_OK; }

int fb200_peer_export(void* d_ptr, unsigned char* handle) {
    static_assert(sizeof(cudaIpcMemHandle_t) == FB200_PEER_HANDLE_BYTES, "handle size");
    if (!d_ptr || !handle) return fail(FB200_EINVAL, "peer_export: bad arguments%s");
    cudaIpcMemHandle_t h;
    int rc = check(cudaIpcGetMemHandle(&h, d_ptr), "peer_export");
    if (rc == FB200_OK) memcpy(handle, &h, sizeof(h));
    return rc;
}

int fb200_peer_open(const unsigned char* handle, void** d_ptr) {
    if (!d_ptr || !handle) return fail(FB200_EINVAL, "peer_open: bad arguments%s");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    return check(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess), "peer_open");
}

int fb200_peer_close(void* d_ptr) { return d_ptr ? check(cudaIpcCloseMemHandle(d_ptr), "peer_close") : FB200_OK; }

}  // extern "C"
