/* Prime-field template (ORACLE -- test infrastructure only, never linked into the product).
 * Include with:  FP = name prefix, NL = number of 64-bit limbs, FP_MOD/FP_R1/FP_R2/FP_INV constants.
 * Restates what arkworks' `Fp<MontBackend<_, N>>` (ark-ff 0.4.2, not vendored; pinned by
 * /root/reference/Cargo.toml:33-40) computes: N little-endian u64 limbs in Montgomery form. */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(n) CAT(FP, n)

typedef struct { uint64_t l[NL]; } FN(_t);

static const FN(_t) FN(_mod) = {FP_MOD};
static const FN(_t) FN(_one) = {FP_R1};
static const FN(_t) FN(_r2) = {FP_R2};

static inline int FN(_is_zero)(const FN(_t) * a) {
  uint64_t o = 0;
  for (int i = 0; i < NL; i++) o |= a->l[i];
  return o == 0;
}
static inline int FN(_eq)(const FN(_t) * a, const FN(_t) * b) {
  uint64_t o = 0;
  for (int i = 0; i < NL; i++) o |= a->l[i] ^ b->l[i];
  return o == 0;
}
static inline int FN(_geq_mod)(const uint64_t *a) {
  for (int i = NL - 1; i >= 0; i--) {
    if (a[i] > FN(_mod).l[i]) return 1;
    if (a[i] < FN(_mod).l[i]) return 0;
  }
  return 1;
}
static inline void FN(_sub_mod_raw)(uint64_t *a) {
  unsigned __int128 br = 0;
  for (int i = 0; i < NL; i++) {
    unsigned __int128 d = (unsigned __int128)a[i] - FN(_mod).l[i] - (uint64_t)br;
    a[i] = (uint64_t)d;
    br = (d >> 64) & 1;
  }
}
static inline void FN(_add)(FN(_t) * r, const FN(_t) * a, const FN(_t) * b) {
  unsigned __int128 c = 0;
  uint64_t t[NL];
  for (int i = 0; i < NL; i++) {
    c += (unsigned __int128)a->l[i] + b->l[i];
    t[i] = (uint64_t)c;
    c >>= 64;
  }
  if (c || FN(_geq_mod)(t)) FN(_sub_mod_raw)(t);
  for (int i = 0; i < NL; i++) r->l[i] = t[i];
}
static inline void FN(_sub)(FN(_t) * r, const FN(_t) * a, const FN(_t) * b) {
  unsigned __int128 br = 0;
  uint64_t t[NL];
  for (int i = 0; i < NL; i++) {
    unsigned __int128 d = (unsigned __int128)a->l[i] - b->l[i] - (uint64_t)br;
    t[i] = (uint64_t)d;
    br = (d >> 64) & 1;
  }
  if (br) {
    unsigned __int128 c = 0;
    for (int i = 0; i < NL; i++) {
      c += (unsigned __int128)t[i] + FN(_mod).l[i];
      t[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  for (int i = 0; i < NL; i++) r->l[i] = t[i];
}
static inline void FN(_neg)(FN(_t) * r, const FN(_t) * a) {
  if (FN(_is_zero)(a)) { *r = *a; return; }
  FN(_t) z;
  for (int i = 0; i < NL; i++) z.l[i] = 0;
  FN(_sub)(r, &z, a);
}
static inline void FN(_dbl)(FN(_t) * r, const FN(_t) * a) { FN(_add)(r, a, a); }

/* Montgomery product a*b*R^-1 (CIOS). */
static inline void FN(_mul)(FN(_t) * r, const FN(_t) * a, const FN(_t) * b) {
  uint64_t t[NL + 2];
  for (int i = 0; i < NL + 2; i++) t[i] = 0;
  for (int i = 0; i < NL; i++) {
    unsigned __int128 c = 0;
    for (int j = 0; j < NL; j++) {
      c += (unsigned __int128)a->l[j] * b->l[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[NL];
    t[NL] = (uint64_t)c;
    t[NL + 1] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * (uint64_t)FP_INV;
    c = (unsigned __int128)m * FN(_mod).l[0] + t[0];
    c >>= 64;
    for (int j = 1; j < NL; j++) {
      c += (unsigned __int128)m * FN(_mod).l[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[NL];
    t[NL - 1] = (uint64_t)c;
    t[NL] = t[NL + 1] + (uint64_t)(c >> 64);
  }
  if (t[NL] || FN(_geq_mod)(t)) FN(_sub_mod_raw)(t);
  for (int i = 0; i < NL; i++) r->l[i] = t[i];
}
static inline void FN(_sqr)(FN(_t) * r, const FN(_t) * a) { FN(_mul)(r, a, a); }

static inline void FN(_to_mont)(FN(_t) * r, const FN(_t) * a) { FN(_mul)(r, a, &FN(_r2)); }
static inline void FN(_from_mont)(FN(_t) * r, const FN(_t) * a) {
  FN(_t) one;
  for (int i = 0; i < NL; i++) one.l[i] = 0;
  one.l[0] = 1;
  FN(_mul)(r, a, &one);
}
/* a^e, e given as nl little-endian limbs (plain integer). */
static inline void FN(_pow)(FN(_t) * r, const FN(_t) * a, const uint64_t *e, int nl) {
  FN(_t) acc = FN(_one);
  for (int i = nl * 64 - 1; i >= 0; i--) {
    FN(_sqr)(&acc, &acc);
    if ((e[i / 64] >> (i % 64)) & 1) FN(_mul)(&acc, &acc, a);
  }
  *r = acc;
}
static inline void FN(_inv)(FN(_t) * r, const FN(_t) * a) {
  uint64_t e[NL];
  for (int i = 0; i < NL; i++) e[i] = FN(_mod).l[i];
  e[0] -= 2; /* moduli are odd and > 2: no borrow */
  FN(_pow)(r, a, e, NL);
}

#undef FN
#undef FP
#undef NL
#undef FP_MOD
#undef FP_R1
#undef FP_R2
#undef FP_INV
