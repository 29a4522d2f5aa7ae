/* Short-Weierstrass (a = 0) group template over a field EF (ORACLE -- test infrastructure only).
 * Include with EC = group prefix, EF = coordinate-field prefix.
 * Restates the semantics of arkworks' `short_weierstrass::{Affine, Projective}` (Jacobian) and
 * `VariableBaseMSM::msm_unchecked` (ark-ec 0.4.2, not vendored) as called from
 * /root/reference/mpc-core/src/protocols/{plain.rs:408-416, rep3.rs:934-947, shamir.rs:1027-1039}.
 * Affine points are packed (x, y) in Montgomery form with (0, 0) = infinity, i.e. the zkey layout
 * (/root/reference/co-circom/circom-types/src/traits.rs:107-155). */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define ECN(n) CAT(EC, n)
#define EFN(n) CAT(EF, n)
typedef EFN(_t) ECN(_f);

typedef struct { ECN(_f) x, y; } ECN(_aff);
typedef struct { ECN(_f) x, y, z; } ECN(_jac);

static inline int ECN(_aff_is_inf)(const ECN(_aff) * p) { return EFN(_is_zero)(&p->x) && EFN(_is_zero)(&p->y); }
static inline int ECN(_is_inf)(const ECN(_jac) * p) { return EFN(_is_zero)(&p->z); }
static inline void ECN(_set_inf)(ECN(_jac) * p) {
  memset(p, 0, sizeof(*p));
  p->x = EFN(_one);
  p->y = EFN(_one);
}
static inline void ECN(_from_aff)(ECN(_jac) * r, const ECN(_aff) * p) {
  if (ECN(_aff_is_inf)(p)) { ECN(_set_inf)(r); return; }
  r->x = p->x;
  r->y = p->y;
  r->z = EFN(_one);
}
static inline void ECN(_dbl)(ECN(_jac) * r, const ECN(_jac) * p) {
  if (ECN(_is_inf)(p)) { *r = *p; return; }
  ECN(_f) A, B, C, D, E, F, t;
  EFN(_sqr)(&A, &p->x);
  EFN(_sqr)(&B, &p->y);
  EFN(_sqr)(&C, &B);
  EFN(_add)(&t, &p->x, &B);
  EFN(_sqr)(&t, &t);
  EFN(_sub)(&t, &t, &A);
  EFN(_sub)(&t, &t, &C);
  EFN(_dbl)(&D, &t);
  EFN(_dbl)(&E, &A);
  EFN(_add)(&E, &E, &A);
  EFN(_sqr)(&F, &E);
  ECN(_f) z3;
  EFN(_mul)(&z3, &p->y, &p->z);
  EFN(_dbl)(&z3, &z3);
  EFN(_dbl)(&t, &D);
  EFN(_sub)(&r->x, &F, &t);
  EFN(_sub)(&t, &D, &r->x);
  EFN(_mul)(&t, &E, &t);
  EFN(_dbl)(&C, &C);
  EFN(_dbl)(&C, &C);
  EFN(_dbl)(&C, &C);
  EFN(_sub)(&r->y, &t, &C);
  r->z = z3;
}
/* r = p + q (q affine) */
static inline void ECN(_madd)(ECN(_jac) * r, const ECN(_jac) * p, const ECN(_aff) * q) {
  if (ECN(_aff_is_inf)(q)) { *r = *p; return; }
  if (ECN(_is_inf)(p)) { ECN(_from_aff)(r, q); return; }
  ECN(_f) Z1Z1, U2, S2, H, HH, I, J, rr, V, t;
  EFN(_sqr)(&Z1Z1, &p->z);
  EFN(_mul)(&U2, &q->x, &Z1Z1);
  EFN(_mul)(&S2, &q->y, &p->z);
  EFN(_mul)(&S2, &S2, &Z1Z1);
  EFN(_sub)(&H, &U2, &p->x);
  EFN(_sub)(&rr, &S2, &p->y);
  if (EFN(_is_zero)(&H)) {
    if (EFN(_is_zero)(&rr)) { ECN(_dbl)(r, p); return; }
    ECN(_set_inf)(r);
    return;
  }
  EFN(_dbl)(&rr, &rr);
  EFN(_sqr)(&HH, &H);
  EFN(_dbl)(&I, &HH);
  EFN(_dbl)(&I, &I);
  EFN(_mul)(&J, &H, &I);
  EFN(_mul)(&V, &p->x, &I);
  ECN(_f) x3, y3, z3;
  EFN(_sqr)(&x3, &rr);
  EFN(_sub)(&x3, &x3, &J);
  EFN(_sub)(&x3, &x3, &V);
  EFN(_sub)(&x3, &x3, &V);
  EFN(_sub)(&t, &V, &x3);
  EFN(_mul)(&y3, &rr, &t);
  EFN(_mul)(&t, &p->y, &J);
  EFN(_dbl)(&t, &t);
  EFN(_sub)(&y3, &y3, &t);
  EFN(_add)(&z3, &p->z, &H);
  EFN(_sqr)(&z3, &z3);
  EFN(_sub)(&z3, &z3, &Z1Z1);
  EFN(_sub)(&z3, &z3, &HH);
  r->x = x3;
  r->y = y3;
  r->z = z3;
}
static inline void ECN(_add)(ECN(_jac) * r, const ECN(_jac) * p, const ECN(_jac) * q) {
  if (ECN(_is_inf)(p)) { *r = *q; return; }
  if (ECN(_is_inf)(q)) { *r = *p; return; }
  ECN(_f) Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t;
  EFN(_sqr)(&Z1Z1, &p->z);
  EFN(_sqr)(&Z2Z2, &q->z);
  EFN(_mul)(&U1, &p->x, &Z2Z2);
  EFN(_mul)(&U2, &q->x, &Z1Z1);
  EFN(_mul)(&S1, &p->y, &q->z);
  EFN(_mul)(&S1, &S1, &Z2Z2);
  EFN(_mul)(&S2, &q->y, &p->z);
  EFN(_mul)(&S2, &S2, &Z1Z1);
  EFN(_sub)(&H, &U2, &U1);
  EFN(_sub)(&rr, &S2, &S1);
  if (EFN(_is_zero)(&H)) {
    if (EFN(_is_zero)(&rr)) { ECN(_dbl)(r, p); return; }
    ECN(_set_inf)(r);
    return;
  }
  EFN(_dbl)(&rr, &rr);
  EFN(_dbl)(&I, &H);
  EFN(_sqr)(&I, &I);
  EFN(_mul)(&J, &H, &I);
  EFN(_mul)(&V, &U1, &I);
  ECN(_f) x3, y3, z3;
  EFN(_sqr)(&x3, &rr);
  EFN(_sub)(&x3, &x3, &J);
  EFN(_sub)(&x3, &x3, &V);
  EFN(_sub)(&x3, &x3, &V);
  EFN(_sub)(&t, &V, &x3);
  EFN(_mul)(&y3, &rr, &t);
  EFN(_mul)(&t, &S1, &J);
  EFN(_dbl)(&t, &t);
  EFN(_sub)(&y3, &y3, &t);
  EFN(_add)(&z3, &p->z, &q->z);
  EFN(_sqr)(&z3, &z3);
  EFN(_sub)(&z3, &z3, &Z1Z1);
  EFN(_sub)(&z3, &z3, &Z2Z2);
  EFN(_mul)(&z3, &z3, &H);
  r->x = x3;
  r->y = y3;
  r->z = z3;
}
static inline void ECN(_neg)(ECN(_jac) * r, const ECN(_jac) * p) {
  *r = *p;
  EFN(_neg)(&r->y, &p->y);
}
static inline void ECN(_to_aff)(ECN(_aff) * r, const ECN(_jac) * p) {
  if (ECN(_is_inf)(p)) { memset(r, 0, sizeof(*r)); return; }
  ECN(_f) zi, zi2;
  EFN(_inv)(&zi, &p->z);
  EFN(_sqr)(&zi2, &zi);
  EFN(_mul)(&r->x, &p->x, &zi2);
  EFN(_mul)(&zi2, &zi2, &zi);
  EFN(_mul)(&r->y, &p->y, &zi2);
}
/* r = k*p, k a plain integer of 4 limbs */
static inline void ECN(_mul)(ECN(_jac) * r, const ECN(_jac) * p, const uint64_t *k) {
  ECN(_jac) acc;
  ECN(_set_inf)(&acc);
  for (int i = 255; i >= 0; i--) {
    ECN(_dbl)(&acc, &acc);
    if ((k[i / 64] >> (i % 64)) & 1) ECN(_add)(&acc, &acc, p);
  }
  *r = acc;
}

/* Signed-digit bucket (Pippenger) MSM over pre-recoded digits (see signed_digits() in
 * cocg_oracle.c).  Window rule as ark-ec 0.4.2 (c = 3 if n < 32 else floor(log2(n)*69/100)+2;
 * quoted from memory).  One call = one window; the caller runs the
 * windows in parallel with OpenMP, like arkworks' rayon path. */
static void ECN(_msm_one_window)(const ECN(_aff) * pts, const int32_t *dig, size_t n, int c, int w, int nwin, ECN(_jac) * out) {
  size_t nb = (size_t)1 << (c - 1);
  ECN(_jac) *bk = (ECN(_jac) *)malloc(nb * sizeof(ECN(_jac)));
  for (size_t i = 0; i < nb; i++) ECN(_set_inf)(&bk[i]);
  for (size_t i = 0; i < n; i++) {
    int32_t d = dig[i * (size_t)nwin + w];
    if (d == 0) continue;
    if (d > 0) ECN(_madd)(&bk[d - 1], &bk[d - 1], &pts[i]);
    else {
      ECN(_aff) np = pts[i];
      EFN(_neg)(&np.y, &np.y);
      ECN(_madd)(&bk[-d - 1], &bk[-d - 1], &np);
    }
  }
  ECN(_jac) run, acc;
  ECN(_set_inf)(&run);
  ECN(_set_inf)(&acc);
  for (size_t i = nb; i-- > 0;) {
    ECN(_add)(&run, &run, &bk[i]);
    ECN(_add)(&acc, &acc, &run);
  }
  free(bk);
  *out = acc;
}
/* Fold window sums: sum_w 2^(c*w) * win[w]. */
static void ECN(_msm_fold)(const ECN(_jac) * win, int nwin, int c, ECN(_jac) * out) {
  ECN(_jac) acc = win[nwin - 1];
  for (int w = nwin - 2; w >= 0; w--) {
    for (int k = 0; k < c; k++) ECN(_dbl)(&acc, &acc);
    ECN(_add)(&acc, &acc, &win[w]);
  }
  *out = acc;
}

#undef ECN
#undef EFN
#undef EC
#undef EF
