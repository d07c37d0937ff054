//
// motion_bounds.cpp -- interpolator segments and motion bounding boxes of animated assembly
// instances (see motion_bounds.h for the reference map).  Plain double arithmetic, written out in
// the reference's evaluation order; compiled with -ffp-contract=off like everything on the host.
//
#include "motion_bounds.h"

#include <algorithm>
#include <cmath>
#include <limits>

namespace asgpu
{

namespace
{

struct V3 { double x, y, z; };
struct Quat { double s; V3 v; };
struct Box { V3 lo, hi; };

// foundation/math/vector.h: dot() starts from 0 and adds the products left to right.
inline double dot3(const V3& a, const V3& b)
{
    double r = 0.0;
    r += a.x * b.x;
    r += a.y * b.y;
    r += a.z * b.z;
    return r;
}

inline double length3(const V3& a) { return std::sqrt(dot3(a, a)); }

// vector.h:639-642, 713-716: a double vector divided by a scalar is multiplied by its reciprocal.
inline V3 over(const V3& a, const double d)
{
    const double r = 1.0 / d;
    return V3{ a.x * r, a.y * r, a.z * r };
}

inline V3 cross3(const V3& l, const V3& r)
{
    return V3{ l.y * r.z - r.y * l.z, l.z * r.x - r.z * l.x, l.x * r.y - r.x * l.y };
}

// Matrix<T, 3, 3>::extract_unit_quaternion (matrix.h:1432-1484, after Shoemake / Eberly).
Quat rotation_to_quaternion(const double m[9])
{
    Quat q;
    const double trace = m[0] + m[4] + m[8];
    if (trace > 0.0)
    {
        double root = std::sqrt(trace + 1.0);
        q.s = 0.5 * root;
        root = 0.5 / root;
        q.v.x = (m[7] - m[5]) * root;
        q.v.y = (m[2] - m[6]) * root;
        q.v.z = (m[3] - m[1]) * root;
        return q;
    }
    size_t i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[i * 3 + i]) i = 2;
    const size_t j = (size_t(1) << i) & 3;
    const size_t k = (size_t(1) << j) & 3;
    double root = std::sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
    double v[3];
    v[i] = 0.5 * root;
    root = 0.5 / root;
    q.s = (m[k * 3 + j] - m[j * 3 + k]) * root;
    v[j] = (m[j * 3 + i] + m[i * 3 + j]) * root;
    v[k] = (m[k * 3 + i] + m[i * 3 + k]) * root;
    q.v = V3{ v[0], v[1], v[2] };
    return q;
}

// Matrix<T, 4, 4>::decompose (matrix.h:2129-2137) of a row-major local-to-parent matrix.
void decompose(const double m[16], V3& scaling, Quat& rotation, V3& translation)
{
    double r[9] = { m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10] };
    scaling.x = length3(V3{ r[0], r[3], r[6] });
    scaling.y = length3(V3{ r[1], r[4], r[7] });
    scaling.z = length3(V3{ r[2], r[5], r[8] });
    const double determinant =
        r[0] * r[4] * r[8] + r[1] * r[5] * r[6] + r[2] * r[3] * r[7] - r[2] * r[4] * r[6] - r[1] * r[3] * r[8] - r[0] * r[5] * r[7];
    if (determinant < 0.0) scaling.x = -scaling.x;
    const double rx = 1.0 / scaling.x, ry = 1.0 / scaling.y, rz = 1.0 / scaling.z;
    r[0] *= rx; r[3] *= rx; r[6] *= rx;
    r[1] *= ry; r[4] *= ry; r[7] *= ry;
    r[2] *= rz; r[5] *= rz; r[8] *= rz;
    rotation = rotation_to_quaternion(r);
    translation = V3{ m[3], m[7], m[11] };
}

inline double quat_dot(const Quat& a, const Quat& b) { return a.s * b.s + dot3(a.v, b.v); }

// feq(lhs, 1.0, eps) of foundation/math/scalar.h:973-1000.
bool near_one(const double lhs, const double eps)
{
    if (lhs == 0.0) return std::abs(1.0) < eps;
    const double ratio = lhs / 1.0;
    return ratio >= 1.0 - eps && ratio <= 1.0 + eps;
}

struct Segment { V3 s0, s1, t0, t1; Quat q0, q1; bool valid; };

// TransformInterpolator::set_transforms (transform.h:641-655).
Segment make_segment(const double from[16], const double to[16])
{
    Segment g;
    decompose(from, g.s0, g.q0, g.t0);
    decompose(to, g.s1, g.q1, g.t1);
    if (quat_dot(g.q0, g.q1) < 0.0)
    {
        g.q1.s = -g.q1.s;
        g.q1.v = V3{ -g.q1.v.x, -g.q1.v.y, -g.q1.v.z };
    }
    g.valid = near_one(quat_dot(g.q0, g.q0), 1.0e-6) && near_one(quat_dot(g.q1, g.q1), 1.0e-6);
    return g;
}

// Transform<double>::point_to_parent<double> / point_to_local<double> (transform.h:311-375): the
// full 4 x 4 product, homogeneous divide only when w != 1.
V3 transform_point(const double m[16], const V3& p)
{
    V3 r;
    r.x = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    r.y = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    r.z = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    const double w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    if (w != 1.0)
    {
        const double rcp = 1.0 / w;
        r.x *= rcp; r.y *= rcp; r.z *= rcp;
    }
    return r;
}

inline void insert(Box& b, const V3& p)
{
    if (b.lo.x > p.x) b.lo.x = p.x;
    if (b.hi.x < p.x) b.hi.x = p.x;
    if (b.lo.y > p.y) b.lo.y = p.y;
    if (b.hi.y < p.y) b.hi.y = p.y;
    if (b.lo.z > p.z) b.lo.z = p.z;
    if (b.hi.z < p.z) b.hi.z = p.z;
}

inline Box empty_box()
{
    const double hi = std::numeric_limits<double>::max();
    return Box{ V3{ hi, hi, hi }, V3{ -hi, -hi, -hi } };
}

// Transform::to_parent(AABB) (transform.h:528-546), double flavour.
Box box_to_parent(const double m[16], const Box& b)
{
    Box r = empty_box();
    insert(r, transform_point(m, V3{ b.lo.x, b.lo.y, b.lo.z }));
    insert(r, transform_point(m, V3{ b.lo.x, b.lo.y, b.hi.z }));
    insert(r, transform_point(m, V3{ b.lo.x, b.hi.y, b.hi.z }));
    insert(r, transform_point(m, V3{ b.lo.x, b.hi.y, b.lo.z }));
    insert(r, transform_point(m, V3{ b.hi.x, b.hi.y, b.lo.z }));
    insert(r, transform_point(m, V3{ b.hi.x, b.hi.y, b.hi.z }));
    insert(r, transform_point(m, V3{ b.hi.x, b.lo.y, b.hi.z }));
    insert(r, transform_point(m, V3{ b.hi.x, b.lo.y, b.lo.z }));
    return r;
}

// Matrix<T, 4, 4>::make_rotation(axis, cos, sin) (matrix.h: rotation about a unit axis).
void rotation_about(const V3& a, const double c, const double s, double m[16])
{
    const double k = 1.0 - c;
    m[0] = k * a.x * a.x + c;        m[1] = k * a.x * a.y - s * a.z;  m[2] = k * a.x * a.z + s * a.y;   m[3] = 0.0;
    m[4] = k * a.y * a.x + s * a.z;  m[5] = k * a.y * a.y + c;        m[6] = k * a.y * a.z - s * a.x;   m[7] = 0.0;
    m[8] = k * a.z * a.x - s * a.y;  m[9] = k * a.z * a.y + s * a.x;  m[10] = k * a.z * a.z + c;        m[11] = 0.0;
    m[12] = 0.0; m[13] = 0.0; m[14] = 0.0; m[15] = 1.0;
}

// The scaling of one axis as a linear function of the rotation angle (transformsequence.cpp:361-384).
struct Ramp
{
    double from, to, rcp_span, slope;
    Ramp(const double a, const double b, const double span) : from(a), to(b), rcp_span(1.0 / span), slope((b - a) * (1.0 / span)) {}
    double at(const double theta) const { const double x = theta * rcp_span; return (1.0 - x) * from + x * to; }
};

// Trajectory of a corner (px, py) in the plane perpendicular to the rotation axis
// (transformsequence.cpp:386-456): value, first and second derivative with respect to the angle.
struct PathX
{
    const Ramp& sx; const Ramp& sy; double px, py;
    double f(const double t) const { return sx.at(t) * std::cos(t) * px - sy.at(t) * std::sin(t) * py; }
    double d(const double t) const
    {
        return (sx.slope * px - sy.at(t) * py) * std::cos(t) - (sx.at(t) * px + sy.slope * py) * std::sin(t);
    }
    double dd(const double t) const
    {
        const double a = sx.slope * px - sy.at(t) * py;
        const double b = sx.at(t) * px + sy.slope * py;
        const double ap = -sy.slope * py;
        const double bp = sx.slope * px;
        return (ap - b) * std::cos(t) - (bp + a) * std::sin(t);
    }
};

struct PathY
{
    const Ramp& sx; const Ramp& sy; double px, py;
    double f(const double t) const { return sx.at(t) * std::sin(t) * px + sy.at(t) * std::cos(t) * py; }
    double d(const double t) const
    {
        return (sx.at(t) * px + sy.slope * py) * std::cos(t) + (sx.slope * px - sy.at(t) * py) * std::sin(t);
    }
    double dd(const double t) const
    {
        const double a = sx.at(t) * px + sy.slope * py;
        const double b = sx.slope * px - sy.at(t) * py;
        const double ap = sx.slope * px;
        const double bp = -sy.slope * py;
        return (ap + b) * std::cos(t) + (bp - a) * std::sin(t);
    }
};

// find_root_newton + find_multiple_roots_newton (foundation/math/root.h:136-219) on the derivative
// of a trajectory: every root found is handed to `found`.
template <typename Path, typename Found>
void extrema(const Path& path, const double a, const double b, const double max_length, const double eps, const size_t max_iterations, Found& found)
{
    if (b - a > max_length)
    {
        const double m = (a + b) * 0.5;
        extrema(path, a, m, max_length, eps, max_iterations, found);
        extrema(path, m, b, max_length, eps, max_iterations, found);
        return;
    }
    const double fa = path.d(a), fb = path.d(b);
    if (fa * fb > 0.0) return;
    double root = (a + b) * 0.5;
    for (size_t i = 0; i < max_iterations; ++i)
    {
        const double delta = path.d(root) / path.dd(root);
        root -= delta;
        if (root < a || root > b) root = root < a ? a : (root > b ? b : root);
        else if (std::abs(delta) <= eps) { found(root); return; }
    }
}

// AABB::robust_grow (aabb.h:621-641).
void robust_grow(Box& b, const double eps)
{
    double* lo[3] = { &b.lo.x, &b.lo.y, &b.lo.z };
    double* hi[3] = { &b.hi.x, &b.hi.y, &b.hi.z };
    double centre[3], extent[3];
    for (int a = 0; a < 3; ++a) { centre[a] = 0.5 * (*lo[a] + *hi[a]); extent[a] = *hi[a] - *lo[a]; }
    for (int a = 0; a < 3; ++a)
    {
        const double dominant = std::max(std::max(std::abs(centre[a]), extent[a]), 1.0);
        const double delta = dominant * eps;
        *lo[a] -= delta;
        *hi[a] += delta;
    }
}

// TransformSequence::compute_motion_segment_bbox (transformsequence.cpp:509-616).
Box segment_bounds(const Box& bbox, const double from_l2p[16], const double to_l2p[16])
{
    const double MinLength = 1.5707963267948966;        // HalfPi<double>()
    const double RootEps = 1.0e-6;
    const double GrowEps = 1.0e-4;
    const size_t MaxIterations = 100;

    const Box from_box = box_to_parent(from_l2p, bbox);
    Box motion = from_box;

    const Segment g = make_segment(from_l2p, to_l2p);
    if (!g.valid) return motion;

    // Relative rotation q1 * conjugate(q0) (quaternion.h: product of two quaternions).
    const Quat c0 = Quat{ g.q0.s, V3{ -g.q0.v.x, -g.q0.v.y, -g.q0.v.z } };
    Quat q;
    q.s = g.q1.s * c0.s - dot3(g.q1.v, c0.v);
    {
        const V3 x = cross3(g.q1.v, c0.v);
        q.v.x = g.q1.s * c0.v.x + c0.s * g.q1.v.x + x.x;
        q.v.y = g.q1.s * c0.v.y + c0.s * g.q1.v.y + x.y;
        q.v.z = g.q1.s * c0.v.z + c0.s * g.q1.v.z + x.z;
    }

    // Quaternion::extract_axis_angle (quaternion.h:259-275).
    V3 axis;
    double angle;
    if (q.s < -1.0 || q.s > 1.0) { angle = 0.0; axis = V3{ 1.0, 0.0, 0.0 }; }
    else
    {
        angle = 2.0 * std::acos(q.s);
        axis = q.v;
        const double n = length3(axis);
        if (n > 0.0) axis = over(axis, n);
        else axis.x = 1.0;
    }
    if (axis.z < 0.0) angle = -angle;
    if (angle == 0.0) return motion;

    // Rotation that takes the rotation axis to Z.
    double to_z[16], from_z[16];
    const V3 perp = cross3(V3{ 0.0, 0.0, 1.0 }, axis);
    const double perp_norm = length3(perp);
    if (perp_norm == 0.0)
    {
        for (int k = 0; k < 16; ++k) to_z[k] = from_z[k] = (k % 5 == 0) ? 1.0 : 0.0;
    }
    else
    {
        const V3 v = over(perp, perp_norm);
        const double sin_a = perp_norm < -1.0 ? -1.0 : (perp_norm > 1.0 ? 1.0 : perp_norm);
        const double cos_a = std::sqrt(1.0 - sin_a * sin_a);
        rotation_about(v, cos_a, +sin_a, from_z);       // axis_to_z.local_to_parent
        rotation_about(v, cos_a, -sin_a, to_z);         // axis_to_z.parent_to_local
    }

    const Ramp sx(1.0, g.s1.x / g.s0.x, angle);
    const Ramp sy(1.0, g.s1.y / g.s0.y, angle);
    const Ramp sz(1.0, g.s1.z / g.s0.z, angle);

    // The four corners at Z = min (AABB::compute_corner(0..3)).
    for (int c = 0; c < 4; ++c)
    {
        const V3 at = V3{ (c & 1) ? from_box.hi.x : from_box.lo.x, (c & 2) ? from_box.hi.y : from_box.lo.y, from_box.lo.z };
        const V3 corner = transform_point(to_z, at);
        const PathX tx{ sx, sy, corner.x, corner.y };
        const PathY ty{ sx, sy, corner.x, corner.y };
        const double a = std::min(angle, 0.0);
        const double b = std::max(angle, 0.0);
        auto found = [&](const double theta)
        {
            const V3 extremum = V3{ tx.f(theta), ty.f(theta), sz.at(theta) * corner.z };
            insert(motion, transform_point(from_z, extremum));
        };
        extrema(tx, a, b, MinLength, RootEps, MaxIterations, found);
        extrema(ty, a, b, MinLength, RootEps, MaxIterations, found);
    }

    robust_grow(motion, GrowEps);
    return motion;
}

// Transform<double>::point_to_parent<float> and to_parent(AABB<float>) (transform.h:346-375, 528-546).
void box_to_parent_float(const double m[16], const float lo[3], const float hi[3], float out_lo[3], float out_hi[3])
{
    for (int a = 0; a < 3; ++a) { out_lo[a] = std::numeric_limits<float>::max(); out_hi[a] = -std::numeric_limits<float>::max(); }
    const int order[8][3] = { {0,0,0}, {0,0,1}, {0,1,1}, {0,1,0}, {1,1,0}, {1,1,1}, {1,0,1}, {1,0,0} };
    for (int k = 0; k < 8; ++k)
    {
        const double px = order[k][0] ? hi[0] : lo[0], py = order[k][1] ? hi[1] : lo[1], pz = order[k][2] ? hi[2] : lo[2];
        float r[3];
        r[0] = static_cast<float>(m[0] * px + m[1] * py + m[2] * pz + m[3]);
        r[1] = static_cast<float>(m[4] * px + m[5] * py + m[6] * pz + m[7]);
        r[2] = static_cast<float>(m[8] * px + m[9] * py + m[10] * pz + m[11]);
        const float w = static_cast<float>(m[12] * px + m[13] * py + m[14] * pz + m[15]);
        if (w != 1.0f)
        {
            const float rcp = 1.0f / w;
            r[0] *= rcp; r[1] *= rcp; r[2] *= rcp;
        }
        for (int a = 0; a < 3; ++a)
        {
            if (out_lo[a] > r[a]) out_lo[a] = r[a];
            if (out_hi[a] < r[a]) out_hi[a] = r[a];
        }
    }
}

}   // anonymous namespace

bool make_transform_segment(const double from_local_to_parent[16], const double to_local_to_parent[16], asgpu_transform_segment& segment)
{
    const Segment g = make_segment(from_local_to_parent, to_local_to_parent);
    segment.s0[0] = g.s0.x; segment.s0[1] = g.s0.y; segment.s0[2] = g.s0.z;
    segment.s1[0] = g.s1.x; segment.s1[1] = g.s1.y; segment.s1[2] = g.s1.z;
    segment.t0[0] = g.t0.x; segment.t0[1] = g.t0.y; segment.t0[2] = g.t0.z;
    segment.t1[0] = g.t1.x; segment.t1[1] = g.t1.y; segment.t1[2] = g.t1.z;
    segment.q0[0] = g.q0.s; segment.q0[1] = g.q0.v.x; segment.q0[2] = g.q0.v.y; segment.q0[3] = g.q0.v.z;
    segment.q1[0] = g.q1.s; segment.q1[1] = g.q1.v.x; segment.q1[2] = g.q1.v.y; segment.q1[3] = g.q1.v.z;
    return g.valid;
}

void motion_bounds(const double* local_to_parent, const double* /*parent_to_local*/, const uint32_t key_count,
                   const float bbox_lo[3], const float bbox_hi[3], float out_lo[3], float out_hi[3])
{
    // TransformSequence::to_parent<float> (transformsequence.h:212-236).
    for (int a = 0; a < 3; ++a) { out_lo[a] = bbox_lo[a]; out_hi[a] = bbox_hi[a]; }
    if (key_count == 0) return;
    for (int a = 0; a < 3; ++a)
        if (!(bbox_lo[a] <= bbox_hi[a])) return;
    for (int a = 0; a < 3; ++a) { out_lo[a] = std::numeric_limits<float>::max(); out_hi[a] = -std::numeric_limits<float>::max(); }
    const Box bbox = Box{ V3{ double(bbox_lo[0]), double(bbox_lo[1]), double(bbox_lo[2]) }, V3{ double(bbox_hi[0]), double(bbox_hi[1]), double(bbox_hi[2]) } };
    for (uint32_t i = 0; i + 1 < key_count; ++i)
    {
        const Box seg = segment_bounds(bbox, local_to_parent + size_t(i) * 16, local_to_parent + size_t(i + 1) * 16);
        const float lo[3] = { static_cast<float>(seg.lo.x), static_cast<float>(seg.lo.y), static_cast<float>(seg.lo.z) };
        const float hi[3] = { static_cast<float>(seg.hi.x), static_cast<float>(seg.hi.y), static_cast<float>(seg.hi.z) };
        for (int a = 0; a < 3; ++a)
        {
            if (out_lo[a] > lo[a]) out_lo[a] = lo[a];
            if (out_hi[a] < hi[a]) out_hi[a] = hi[a];
        }
    }
    float lo[3], hi[3];
    box_to_parent_float(local_to_parent + size_t(key_count - 1) * 16, bbox_lo, bbox_hi, lo, hi);
    for (int a = 0; a < 3; ++a)
    {
        if (out_lo[a] > lo[a]) out_lo[a] = lo[a];
        if (out_hi[a] < hi[a]) out_hi[a] = hi[a];
    }
}

}   // namespace asgpu
