/* Fq2 = Fq[u]/(u^2+1) template (ORACLE -- test infrastructure only).
 * Include with F2 = prefix of the new type, FQ = prefix of the base field. */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define F2N(n) CAT(F2, n)
#define FQN(n) CAT(FQ, n)

typedef struct { FQN(_t) c0, c1; } F2N(_t);

static inline int F2N(_is_zero)(const F2N(_t) * a) { return FQN(_is_zero)(&a->c0) && FQN(_is_zero)(&a->c1); }
static inline int F2N(_eq)(const F2N(_t) * a, const F2N(_t) * b) { return FQN(_eq)(&a->c0, &b->c0) && FQN(_eq)(&a->c1, &b->c1); }
static inline void F2N(_add)(F2N(_t) * r, const F2N(_t) * a, const F2N(_t) * b) {
  FQN(_add)(&r->c0, &a->c0, &b->c0);
  FQN(_add)(&r->c1, &a->c1, &b->c1);
}
static inline void F2N(_sub)(F2N(_t) * r, const F2N(_t) * a, const F2N(_t) * b) {
  FQN(_sub)(&r->c0, &a->c0, &b->c0);
  FQN(_sub)(&r->c1, &a->c1, &b->c1);
}
static inline void F2N(_neg)(F2N(_t) * r, const F2N(_t) * a) {
  FQN(_neg)(&r->c0, &a->c0);
  FQN(_neg)(&r->c1, &a->c1);
}
static inline void F2N(_dbl)(F2N(_t) * r, const F2N(_t) * a) { F2N(_add)(r, a, a); }
static inline void F2N(_mul)(F2N(_t) * r, const F2N(_t) * a, const F2N(_t) * b) {
  FQN(_t) t0, t1, s0, s1, m;
  FQN(_mul)(&t0, &a->c0, &b->c0);
  FQN(_mul)(&t1, &a->c1, &b->c1);
  FQN(_add)(&s0, &a->c0, &a->c1);
  FQN(_add)(&s1, &b->c0, &b->c1);
  FQN(_mul)(&m, &s0, &s1);
  FQN(_sub)(&m, &m, &t0);
  FQN(_sub)(&r->c1, &m, &t1);
  FQN(_sub)(&r->c0, &t0, &t1);
}
static inline void F2N(_sqr)(F2N(_t) * r, const F2N(_t) * a) {
  FQN(_t) s, d, m;
  FQN(_add)(&s, &a->c0, &a->c1);
  FQN(_sub)(&d, &a->c0, &a->c1);
  FQN(_mul)(&m, &a->c0, &a->c1);
  FQN(_mul)(&r->c0, &s, &d);
  FQN(_add)(&r->c1, &m, &m);
}
static inline void F2N(_inv)(F2N(_t) * r, const F2N(_t) * a) {
  FQN(_t) n, t;
  FQN(_sqr)(&n, &a->c0);
  FQN(_sqr)(&t, &a->c1);
  FQN(_add)(&n, &n, &t);
  FQN(_inv)(&n, &n);
  FQN(_mul)(&r->c0, &a->c0, &n);
  FQN(_mul)(&t, &a->c1, &n);
  FQN(_neg)(&r->c1, &t);
}
static const F2N(_t) F2N(_one) = {{FQ_R1_INIT}, {{0}}};

#undef F2N
#undef FQN
#undef F2
#undef FQ
#undef FQ_R1_INIT
