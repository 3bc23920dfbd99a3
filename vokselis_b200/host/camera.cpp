// camera.cpp — host-side orbit camera, the C++ mirror of src/camera.rs (the reference host is Rust,
// which this image cannot build; see INTEGRATION.md for the Rust-side binding).
//
// `Camera` keeps the reference's fields and setters (src/camera.rs:75-157); the matrices follow
// glam 0.20.5 (crates.io, Cargo.lock:564-566): Mat4::look_at_rh, Mat4::perspective_rh (depth 0..1),
// Mat4 * Mat4, Mat4::inverse, column-major storage.
#include "vokselis.hpp"

#include <cmath>
#include <cstring>

namespace vokselis {

namespace {
struct V3 {
    float x, y, z;
};
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline V3 normalized(V3 a) {  // glam: self * length_recip()
    const float r = 1.0f / std::sqrt(dot(a, a));
    return {a.x * r, a.y * r, a.z * r};
}
}  // namespace

Mat4 Mat4::identity() {
    Mat4 m{};
    m.c[0] = m.c[5] = m.c[10] = m.c[15] = 1.0f;
    return m;
}

// glam Mat4::look_at_rh(eye, center, up) == look_to_rh(eye, center - eye, up)
Mat4 Mat4::look_at_rh(const float eye[3], const float center[3], const float up[3]) {
    const V3 e{eye[0], eye[1], eye[2]};
    const V3 f = normalized(sub(V3{center[0], center[1], center[2]}, e));
    const V3 s = normalized(cross(f, V3{up[0], up[1], up[2]}));
    const V3 u = cross(s, f);
    Mat4 m{};
    m.c[0] = s.x; m.c[1] = u.x; m.c[2] = -f.x; m.c[3] = 0.0f;
    m.c[4] = s.y; m.c[5] = u.y; m.c[6] = -f.y; m.c[7] = 0.0f;
    m.c[8] = s.z; m.c[9] = u.z; m.c[10] = -f.z; m.c[11] = 0.0f;
    m.c[12] = -dot(s, e); m.c[13] = -dot(u, e); m.c[14] = dot(f, e); m.c[15] = 1.0f;
    return m;
}

// glam Mat4::perspective_rh: right-handed, depth range [0, 1]
Mat4 Mat4::perspective_rh(float fov_y, float aspect, float z_near, float z_far) {
    const float sf = std::sin(0.5f * fov_y), cf = std::cos(0.5f * fov_y);
    const float h = cf / sf, w = h / aspect, r = z_far / (z_near - z_far);
    Mat4 m{};
    m.c[0] = w;
    m.c[5] = h;
    m.c[10] = r; m.c[11] = -1.0f;
    m.c[14] = r * z_near;
    return m;
}

Mat4 Mat4::operator*(const Mat4& b) const {
    Mat4 o{};
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row) {
            float acc = 0.0f;
            for (int k = 0; k < 4; ++k) acc += c[k * 4 + row] * b.c[col * 4 + k];
            o.c[col * 4 + row] = acc;
        }
    return o;
}

// Gauss-Jordan with partial pivoting in double, rounded once to f32 (glam uses an fp32 cofactor
// expansion; both are well within what a 4x4 proj*view needs, and the 144-B uniform is an INPUT
// of the raycast boundary, so kernel parity never depends on which inverse produced it).
Mat4 Mat4::inverse() const {
    double a[4][8];
    for (int r = 0; r < 4; ++r)
        for (int col = 0; col < 4; ++col) {
            a[r][col] = c[col * 4 + r];
            a[r][4 + col] = r == col ? 1.0 : 0.0;
        }
    for (int i = 0; i < 4; ++i) {
        int piv = i;
        for (int r = i + 1; r < 4; ++r)
            if (std::fabs(a[r][i]) > std::fabs(a[piv][i])) piv = r;
        if (piv != i)
            for (int k = 0; k < 8; ++k) { double t = a[i][k]; a[i][k] = a[piv][k]; a[piv][k] = t; }
        const double d = a[i][i];
        for (int k = 0; k < 8; ++k) a[i][k] /= d;  // singular -> inf/NaN, like glam's unchecked inverse
        for (int r = 0; r < 4; ++r)
            if (r != i) {
                const double f = a[r][i];
                for (int k = 0; k < 8; ++k) a[r][k] -= f * a[i][k];
            }
    }
    Mat4 o{};
    for (int r = 0; r < 4; ++r)
        for (int col = 0; col < 4; ++col) o.c[col * 4 + r] = (float)a[r][4 + col];
    return o;
}

// ---- Camera (src/camera.rs:87-171) -------------------------------------------------------------
Camera::Camera(float zoom_, float pitch_, float yaw_, const float target_[3], float aspect_)
    : zoom(zoom_), pitch(pitch_), yaw(yaw_), aspect(aspect_) {
    target[0] = target_[0]; target[1] = target_[1]; target[2] = target_[2];
    up[0] = 0.0f; up[1] = 1.0f; up[2] = 0.0f;  // Camera::UP = Vec3::Y (:91)
    updated = false;                           // :103
    fix_eye();
}

void Camera::fix_eye() {  // :148-157
    const float pitch_cos = std::cos(pitch);
    eye[0] = target[0] - zoom * (std::sin(yaw) * pitch_cos);
    eye[1] = target[1] - zoom * std::sin(pitch);
    eye[2] = target[2] - zoom * (std::cos(yaw) * pitch_cos);
}

Mat4 Camera::build_projection_view_matrix() const {  // :109-113
    const Mat4 view = Mat4::look_at_rh(eye, target, up);
    const Mat4 proj = Mat4::perspective_rh(FOVY, aspect, ZNEAR, ZFAR);
    return proj * view;
}

static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

void Camera::set_zoom(float z) {  // :115-119
    zoom = clampf(z, 0.3f, ZFAR / 2.0f);
    fix_eye();
    updated = true;
}
void Camera::add_zoom(float d) { set_zoom(zoom + d); }
void Camera::set_pitch(float p) {  // :125-132
    const float eps = 1.1920929e-07f, half_pi = 3.14159265358979323846f / 2.0f;
    pitch = clampf(p, -half_pi + eps, half_pi - eps);
    fix_eye();
    updated = true;
}
void Camera::add_pitch(float d) { set_pitch(pitch + d); }
void Camera::set_yaw(float y) {  // :138-142
    yaw = y;
    fix_eye();
    updated = true;
}
void Camera::add_yaw(float d) { set_yaw(yaw + d); }
void Camera::set_aspect(unsigned width, unsigned height) {  // :159-162
    aspect = (float)width / (float)height;
    updated = true;
}

VkrtCameraUniform Camera::get_proj_view_matrix() const {  // :164-171
    const Mat4 pv = build_projection_view_matrix();
    const Mat4 inv = pv.inverse();
    VkrtCameraUniform u;
    u.view_position[0] = eye[0]; u.view_position[1] = eye[1]; u.view_position[2] = eye[2]; u.view_position[3] = 1.0f;
    std::memcpy(u.proj_view, pv.c, sizeof pv.c);
    std::memcpy(u.inv_proj, inv.c, sizeof inv.c);
    return u;
}

}  // namespace vokselis

extern "C" int vkrt_camera_uniform(float zoom, float pitch, float yaw, const float target[3], float aspect, VkrtCameraUniform* out) {
    if (!target || !out) return VKRT_ERR_INVALID;
    *out = vokselis::Camera(zoom, pitch, yaw, target, aspect).get_proj_view_matrix();
    return VKRT_OK;
}
