// Association value functions shared by the OC-SORT family kernels (ocsort.cu, deepocsort.cu): boxes x1y1x2y2, float64.
// /root/reference/plugins/track/oc_sort/association.py:5-171 (the deep_oc_sort copy, association.py:9-199, is the same arithmetic).
#pragma once
#include "tk_common.cuh"

namespace tk {

__device__ __forceinline__ double iou_plain(const double* a, const double* b) {
    const double w = fmax(0.0, fmin(a[2], b[2]) - fmax(a[0], b[0]));
    const double h = fmax(0.0, fmin(a[3], b[3]) - fmax(a[1], b[1]));
    const double wh = __dmul_rn(w, h);
    const double ua = __dsub_rn(__dadd_rn(__dmul_rn(a[2] - a[0], a[3] - a[1]), __dmul_rn(b[2] - b[0], b[3] - b[1])), wh);
    return wh / ua;
}

static __device__ double asso_value(int kind, const double* a, const double* b) {
    const double iou = iou_plain(a, b);
    if (kind == 0) return iou;
    const double w = fmax(0.0, fmin(a[2], b[2]) - fmax(a[0], b[0]));
    const double h = fmax(0.0, fmin(a[3], b[3]) - fmax(a[1], b[1]));
    const double wh = __dmul_rn(w, h);
    const double wc = fmax(a[2], b[2]) - fmin(a[0], b[0]);
    const double hc = fmax(a[3], b[3]) - fmin(a[1], b[1]);
    if (kind == 1) {  // giou (association.py:24-55)
        const double hull = __dmul_rn(wc, hc);
        const double g = __dsub_rn(iou, __dsub_rn(hull, wh) / hull);
        return __dadd_rn(g, 1.0) / 2.0;
    }
    const double dcx = __dsub_rn((a[0] + a[2]) / 2.0, (b[0] + b[2]) / 2.0);
    const double dcy = __dsub_rn((a[1] + a[3]) / 2.0, (b[1] + b[3]) / 2.0);
    const double inner = __dadd_rn(__dmul_rn(dcx, dcx), __dmul_rn(dcy, dcy));
    const double outer = __dadd_rn(__dmul_rn(wc, wc), __dmul_rn(hc, hc));
    if (kind == 2) return __dadd_rn(__dsub_rn(iou, inner / outer), 1.0) / 2.0;  // diou (:58-95)
    // ciou (:97-147)
    const double w1 = a[2] - a[0], h1 = (a[3] - a[1]) + 1.0, w2 = b[2] - b[0], h2 = (b[3] - b[1]) + 1.0;
    const double at = __dsub_rn(atan(w2 / h2), atan(w1 / h1));
    const double pi = 3.141592653589793;
    const double v = __dmul_rn(4.0 / __dmul_rn(pi, pi), __dmul_rn(at, at));
    const double alpha = v / __dadd_rn(__dsub_rn(1.0, iou), v);
    return __dadd_rn(__dsub_rn(__dsub_rn(iou, inner / outer), __dmul_rn(alpha, v)), 1.0) / 2.0;
}

// ct_dist (association.py:150-171): centre distance d, then (d / d.max()).max() - d / d.max() = 1 - d / d.max()
// (NaN everywhere when all centres coincide, exactly like the 0/0 of the reference). Two passes over the matrix.
__device__ __forceinline__ double centre_dist(const double* a, const double* b) {
    const double dx = (a[0] + a[2]) / 2.0 - (b[0] + b[2]) / 2.0, dy = (a[1] + a[3]) / 2.0 - (b[1] + b[3]) / 2.0;
    return sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

}  // namespace tk
