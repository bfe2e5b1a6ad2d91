"""CPU restatement of the multi-scale / flip testing merge of the reference.  TEST INFRASTRUCTURE ONLY.

    test_net's per-class gathering over the (scale, flip) passes     reference multi_eval.py:557-640
    bbox_vote (score-weighted box voting)                            reference multi_eval.py:453-494

Everything is NumPy float32 arithmetic exactly as the reference performs it (the detections it stacks are float32 arrays,
`0.45` and the size thresholds are weak Python scalars), so this module IS the arithmetic definition the device kernel
(tdrn_multiscale_vote) is tested against bit for bit, including NumPy's summation orders:
  * `np.sum(acc[:, 0:4], axis=0)`  -> rows added one after the other, starting from row 0;
  * `np.sum(acc[:, -1:])`          -> NumPy's pairwise sum over all members (n < 8: sequential from 0; n <= 128: 8 strided
                                      accumulators combined as ((0+1)+(2+3))+((4+5)+(6+7)), then the tail; larger n split in
                                      halves rounded down to a multiple of 8) -- pinned against np.sum in tests/test_multi_scale.py.
Pin status: `bbox_vote` is PINNED against the reference's own function -- oracle/make_golden_vote.py executes the source of
multi_eval.py:453-494 (extracted with `ast`; the module cannot be imported) on seeded detections and stores its outputs in
tests/golden/bbox_vote.npz; restatement and device kernel reproduce them bit for bit.  The gathering loop (:557-640) is
inline in `test_net`; tests/test_oracle_vs_reference.py executes the source of the WHOLE `test_net` (with real cv2, the
reference's PriorBox / Detect / Cython NMS and a stand-in network) at base size 512 and `multi_scale_merge` reproduces every
voted box of every class and image bit for bit (build container only).
Pinned where the reference leaves the order open: detections are visited by descending score, ties -> lower index
(`argsort()[::-1]` of an unstable sort in the reference).

Deviation, documented: the reference's size-rule table names the key '320_706' although its scale list says 704
(multi_eval.py:21-24 vs :618), so at base size 320 the last scale falls through with a stale index array; the evident
intent (same rule as '512_1216': min side < 32) is what is implemented on both sides.
"""
import numpy as np

VOTE_THRESH = 0.45            # multi_eval.py:472

# (base size, scale) -> (kind, threshold): 'max_gt' keeps boxes whose longer side exceeds thr, 'min_lt' boxes whose shorter side
# is below it (sides measured with the +1 pixel convention), multi_eval.py:574-625
SIZE_RULES = {
    (320, 192): ('max_gt', 32), (512, 320): ('max_gt', 32),
    (320, 320): ('max_gt', 0), (512, 512): ('max_gt', 0),
    (320, 384): ('min_lt', 160), (512, 640): ('min_lt', 160),
    (320, 448): ('min_lt', 128), (320, 512): ('min_lt', 96), (320, 576): ('min_lt', 64),
    (320, 704): ('min_lt', 32), (512, 1216): ('min_lt', 32),
}


def gather_class(passes, cls, w, h, base):
    """passes: list of (scale, flipped, det [C, top_k, 5] float32 = Detect output of that pass for ONE image), in the
    reference's order (scales ascending as listed, unflipped then flipped).  -> [n, 5] float32 (x1, y1, x2, y2, score)
    in pixels of the original image, multi_eval.py:557-640."""
    rows = []
    for scale, flipped, det in passes:
        d = np.asarray(det[cls], dtype=np.float32)
        if d.sum() == 0:                                             # :561-562
            continue
        d = d[d[:, 0] > 0]                                           # :563-564
        boxes = d[:, 1:].copy()
        if flipped:                                                  # :566-571: x1' = 1 - x2, x2' = 1 - x1
            nx0 = np.float32(1) - boxes[:, 0]
            nx2 = np.float32(1) - boxes[:, 2]
            boxes[:, 0], boxes[:, 2] = nx2, nx0
        boxes[:, 0] *= np.float32(w); boxes[:, 2] *= np.float32(w)   # :572-575
        boxes[:, 1] *= np.float32(h); boxes[:, 3] *= np.float32(h)
        kind, thr = SIZE_RULES[(base, scale)]
        sw = boxes[:, 2] - boxes[:, 0] + np.float32(1)
        sh = boxes[:, 3] - boxes[:, 1] + np.float32(1)
        keep = np.maximum(sw, sh) > thr if kind == 'max_gt' else np.minimum(sw, sh) < thr
        idx = np.where(keep)[0]
        if idx.size == 0:
            continue
        rows.append(np.hstack((boxes[idx], d[idx, 0][:, None])).astype(np.float32))
    return np.concatenate(rows, 0) if rows else np.zeros((0, 5), np.float32)


def bbox_vote(det, thresh=VOTE_THRESH):
    """multi_eval.py:453-494.  det [n,5] float32 (x1,y1,x2,y2,score) -> [m,5] float32 (merged rows are float32-valued)."""
    det = np.asarray(det, dtype=np.float32)
    if det.shape[0] <= 1:
        return det
    det = det[np.argsort(-det[:, 4], kind='stable')]
    out = []
    while det.shape[0] > 0:
        area = (det[:, 2] - det[:, 0] + 1) * (det[:, 3] - det[:, 1] + 1)
        xx1 = np.maximum(det[0, 0], det[:, 0]); yy1 = np.maximum(det[0, 1], det[:, 1])
        xx2 = np.minimum(det[0, 2], det[:, 2]); yy2 = np.minimum(det[0, 3], det[:, 3])
        iw = np.maximum(0.0, xx2 - xx1 + 1); ih = np.maximum(0.0, yy2 - yy1 + 1)
        inter = iw * ih
        with np.errstate(divide='ignore', invalid='ignore'):
            o = inter / (area[0] + area[:] - inter)
        members = np.where(o >= thresh)[0]
        if members.size == 0:                                        # NaN overlap (degenerate box): the reference would spin;
            members = np.array([0])                                  # both sides emit the head alone and move on
        acc = det[members, :]
        det = np.delete(det, members, 0)
        if members.shape[0] <= 1:
            out.append(acc.astype(np.float32))
            continue
        acc[:, 0:4] = acc[:, 0:4] * np.tile(acc[:, -1:], (1, 4))
        row = np.zeros((1, 5))
        row[:, 0:4] = np.sum(acc[:, 0:4], axis=0) / np.sum(acc[:, -1:])
        row[:, 4] = np.max(acc[:, 4])
        out.append(row.astype(np.float32))                           # exactly representable: the quotient is a float32
    return np.concatenate(out, 0)


def multi_scale_merge(passes, num_classes, w, h, base):
    """-> list over classes 1..C-1 of voted [m,5] float32 arrays (all_boxes[j][i] of multi_eval.py:641-644)."""
    res = [np.zeros((0, 5), np.float32)]
    for j in range(1, num_classes):
        c = gather_class(passes, j, w, h, base)
        res.append(bbox_vote(c) if c.size else c)
    return res


def np_pairwise_sum_f32(a):
    """NumPy's float32 add.reduce over a 1-D (strided) array, spelled out; the CPU tests pin it against np.sum itself and the
    kernel emulates exactly this order."""
    a = [np.float32(v) for v in a]

    def pw(v):
        n = len(v)
        if n < 8:
            r = np.float32(0)
            for t in v:
                r = np.float32(r + t)
            return r
        if n <= 128:
            r = list(v[:8])
            i = 8
            while i < n - (n % 8):
                for j in range(8):
                    r[j] = np.float32(r[j] + v[i + j])
                i += 8
            res = np.float32(np.float32(np.float32(r[0] + r[1]) + np.float32(r[2] + r[3])) +
                             np.float32(np.float32(r[4] + r[5]) + np.float32(r[6] + r[7])))
            while i < n:
                res = np.float32(res + v[i]); i += 1
            return res
        n2 = n // 2
        n2 -= n2 % 8
        return np.float32(pw(v[:n2]) + pw(v[n2:]))
    return pw(a)
