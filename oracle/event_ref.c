/* CPU oracle (C): two-window t-statistic event detection.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.  Used by
 * tests/ as a second checker and by bench.py's cpu_baseline / --impl reference
 * legs as the timed CPU implementation of the event path ("port").
 *
 * Restates /root/reference/event_detection/event_detector.py:
 *   _add_sample      :85-107     ring of float64 prefix sums
 *   _compute_tstat   :109-147
 *   _detect_peak     :149-187
 *   _create_event    :189-210
 * Pinned: asserted identical to the reference module on seeded signals
 * (tests/golden/event_*.npz, produced by tools/make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/Makefile).
 * FMA contraction MUST stay off: the reference evaluates every product and sum
 * as a separately rounded float64 operation.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define RV_FLT_MIN 1.17549435e-38
#define RV_FLT_MAX 3.40282347e+38

typedef struct {
    int w;
    double thr;
    uint32_t masked_to;
    int32_t pos;
    double val;
    int valid;
} peak_t;

static void peak_clear(peak_t *d) { d->pos = -1; d->val = RV_FLT_MAX; d->valid = 0; }

static int peak_step(peak_t *d, peak_t *mask_target, double v, uint32_t mid, double ph)
{
    if (d->masked_to >= mid) return 0;
    if (d->pos == -1) {
        if (v < d->val) d->val = v;
        else if (v - d->val > ph) { d->val = v; d->pos = (int32_t)mid; }
        return 0;
    }
    if (v > d->val) { d->val = v; d->pos = (int32_t)mid; }
    if (mask_target && d->val > d->thr) {
        mask_target->masked_to = (uint32_t)((int64_t)d->pos + d->w);
        peak_clear(mask_target);
    }
    if (d->val - v > ph && d->val > d->thr) d->valid = 1;
    if (d->valid && (double)((int64_t)mid - (int64_t)d->pos) > (double)d->w / 2.0) {
        d->pos = -1; d->val = v; d->valid = 0;
        return 1;
    }
    return 0;
}

/* ring[] holds S for the latest n' = slot (mod buf); the streaming form is kept
 * here on purpose (the Python oracle uses the absolute-index form), so the two
 * restatements check each other. */
static double tstat(const double *rs, const double *rq, int buf, uint32_t t, uint32_t mid, int w)
{
    if (t <= (uint32_t)(2 * w) || w < 2) return 0.0;
    double wf = (double)w;
    uint32_t i = mid % (uint32_t)buf;
    uint32_t st = (uint32_t)(mid - (uint32_t)w) % (uint32_t)buf;
    uint32_t en = (uint32_t)(mid + (uint32_t)w) % (uint32_t)buf;
    double sum1 = rs[i] - rs[st], sq1 = rq[i] - rq[st];
    double sum2 = rs[en] - rs[i], sq2 = rq[en] - rq[i];
    double m1 = sum1 / wf, m2 = sum2 / wf;
    double var = sq1 / wf - m1 * m1 + sq2 / wf - m2 * m2;
    if (RV_FLT_MIN > var) var = RV_FLT_MIN;          /* Python max(var, eta) */
    return fabs(m2 - m1) / sqrt(var / wf);
}

/* Returns the number of events (may exceed cap; only the first cap are stored),
 * or -1 on bad arguments.  length is stored as the low 32 bits (the reference
 * produces u32-wrapped lengths during warm-up for exotic window pairs). */
int64_t rvo_detect_events(const int32_t *raw, int64_t n, int w1, int w2,
                          double thr1, double thr2, double peak_height,
                          int32_t *start, int32_t *length, double *mean, double *stdv,
                          int64_t cap)
{
    if (w1 <= 0 || w2 <= 0 || n < 0) return -1;
    int buf = 1 + 2 * w2;
    double *rs = (double *)calloc((size_t)buf * 2, sizeof(double));
    if (!rs) return -1;
    double *rq = rs + buf;
    peak_t sh = { w1, thr1, 0, -1, RV_FLT_MAX, 0 };
    peak_t lo = { w2, thr2, 0, -1, RV_FLT_MAX, 0 };
    uint32_t t = 1, ev_st = 0;
    double ev_sum = 0.0, ev_sq = 0.0;
    int64_t count = 0;
    for (int64_t k = 0; k < n; ++k) {
        double s = (double)raw[k];
        uint32_t tm = t % (uint32_t)buf;
        uint32_t pm = tm > 0 ? tm - 1 : (uint32_t)buf - 1;
        rs[tm] = rs[pm] + s;
        rq[tm] = rq[pm] + s * s;
        t += 1;
        uint32_t mid = t - (uint32_t)w2 - 1u;
        double t1 = tstat(rs, rq, buf, t, mid, w1);
        double t2 = tstat(rs, rq, buf, t, mid, w2);
        int p1 = peak_step(&sh, &lo, t1, mid, peak_height);
        int p2 = peak_step(&lo, (w1 == w2) ? &lo : (peak_t *)0, t2, mid, peak_height);
        if (!(p1 || p2)) continue;
        uint32_t en = mid - (uint32_t)w1 + 1u;
        double len = (double)((int64_t)en - (int64_t)ev_st);
        if (len < RV_FLT_MIN) continue;
        uint32_t slot = en % (uint32_t)buf;
        double m = (rs[slot] - ev_sum) / len;
        double var = (rq[slot] - ev_sq) / len - pow(m, 2.0);
        if (RV_FLT_MIN > var) var = RV_FLT_MIN;
        if (count < cap) {
            start[count] = (int32_t)ev_st;
            length[count] = (int32_t)(uint32_t)(int64_t)len;
            mean[count] = m;
            stdv[count] = sqrt(var);
        }
        ++count;
        ev_st = en; ev_sum = rs[slot]; ev_sq = rq[slot];
    }
    free(rs);
    return count;
}
