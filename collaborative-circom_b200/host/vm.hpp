// Batched witness-extension arithmetic: the field opcodes of the reference's MPC VM
//   Add Sub Mul Neg Div      /root/reference/co-circom/circom-mpc-vm/src/mpc_vm.rs:508-546
// dispatched through CircomWitnessExtensionProtocol::{vm_add, vm_sub, vm_mul, vm_neg, vm_div}
//   Rep3VmType::{add, sub, mul, neg, div}   /root/reference/mpc-core/src/protocols/rep3/witness_extension_impl.rs:81-200
//   PlainDriver                             /root/reference/mpc-core/src/protocols/plain.rs:421-445
// executed as SIMD over a BATCH of B independent inputs of the same circuit: a VM value is a vector of B field elements in HBM
// (Public) or a share vector of B elements (Shared), one opcode is one kernel over the batch, and a shared multiplication costs ONE
// network round for the whole batch instead of one per instance -- the only form in which the interpreter's arithmetic can use a GPU
// (one run of the reference's VM is a dependent chain of scalar operations with a round trip per shared Mul).  The Public / Shared
// case analysis per opcode is the reference's; the circom front end (bytecode, control flow, comparisons and bit operations on shared
// values, which need bit decomposition) is out of scope -- a straight-line program over registers stands in for it, which is also what
// an MpcAccelerator function (circom-mpc-vm/src/accelerator.rs:34-40) is handed: values in, values out.
#pragma once
#include "driver.hpp"

namespace cohost {

enum VmOp { VM_ADD = 0, VM_SUB = 1, VM_MUL = 2, VM_NEG = 3, VM_DIV = 4 };
struct VmInstr {
  int op, dst, lhs, rhs;  // registers; rhs ignored by VM_NEG
};

struct VmValue {
  enum Kind { EMPTY = 0, PUBLIC = 1, SHARED = 2 } kind = EMPTY;
  DevVec pub;         // PUBLIC
  FieldShareVec sh;   // SHARED
};

template <class T>
class BatchedWitnessVm {
 public:
  BatchedWitnessVm(T& driver, size_t batch, int n_regs) : driver(driver), batch(batch), regs(n_regs) {}
  ~BatchedWitnessVm() {
    for (auto& v : regs) clear(v);
  }
  T& driver;
  size_t batch;
  std::vector<VmValue> regs;
  size_t network_rounds = 0;  // shared multiplications / inversions executed: one exchange each for the whole batch

  static constexpr int K = T::kComponents;

  void set_public(int r, const void* host) {
    VmValue& v = at(r);
    clear(v);
    v.kind = VmValue::PUBLIC;
    v.pub = driver.upload(host, batch);
  }
  void set_shared(int r, const void* host_a, const void* host_b) {
    VmValue& v = at(r);
    clear(v);
    v.kind = VmValue::SHARED;
    v.sh = driver.share_vec_from_host(host_a, host_b, batch);
  }
  void run(const VmInstr* prog, size_t n) {
    for (size_t i = 0; i < n; i++) step(prog[i]);
    check(driver.ctx, cocg_sync(driver.ctx), "cocg_sync");
  }
  VmValue& at(int r) {
    if (r < 0 || (size_t)r >= regs.size()) throw Error("vm: register out of range");
    return regs[r];
  }

 private:
  void clear(VmValue& v) {
    driver.release(v.pub);
    driver.release(v.sh);
    v.kind = VmValue::EMPTY;
  }
  const DevVec& comp(const FieldShareVec& v, int k) const { return k == 0 ? v.a : v.b; }
  DevVec& comp(FieldShareVec& v, int k) { return k == 0 ? v.a : v.b; }
  void vop(int op, const DevVec& a, const DevVec* b, DevVec& out) { check(driver.ctx, cocg_vec_op(driver.ctx, op, a.p, b ? b->p : nullptr, out.p, batch), "cocg_vec_op"); }
  DevVec pub_op(int op, const DevVec& a, const DevVec* b) {
    DevVec o = driver.alloc(batch);
    vop(op, a, b, o);
    return o;
  }
  FieldShareVec copy(const FieldShareVec& s) {
    FieldShareVec o = driver.alloc_share(batch);
    for (int k = 0; k < K; k++) check(driver.ctx, cocg_d2d(driver.ctx, comp(o, k).p, comp(s, k).p, batch * 32), "cocg_d2d");
    return o;
  }
  // add_with_public (rep3.rs:600-608): the public vector enters the component that carries public addends
  FieldShareVec add_public(const FieldShareVec& s, const DevVec& p, bool negate_share) {
    FieldShareVec o = driver.alloc_share(batch);
    for (int k = 0; k < K; k++) {
      if (negate_share) vop(COCG_OP_NEG, comp(s, k), nullptr, comp(o, k));
      else check(driver.ctx, cocg_d2d(driver.ctx, comp(o, k).p, comp(s, k).p, batch * 32), "cocg_d2d");
    }
    const int pc = driver.pub_comp();
    if (pc >= 0 && pc < K) vop(COCG_OP_ADD, comp(o, pc), &p, comp(o, pc));
    return o;
  }
  FieldShareVec mul_public(const FieldShareVec& s, const DevVec& p) {  // mul_with_public
    FieldShareVec o = driver.alloc_share(batch);
    for (int k = 0; k < K; k++) vop(COCG_OP_MUL, comp(s, k), &p, comp(o, k));
    return o;
  }
  FieldShareVec mul_shared(const FieldShareVec& a, const FieldShareVec& b) {
    network_rounds++;
    return driver.mul_vec(a, b);
  }
  FieldShareVec inv_shared(const FieldShareVec& a) {
    network_rounds++;
    return driver.inv_many(a);
  }

  void step(const VmInstr& in) {
    const VmValue& a = at(in.lhs);
    if (a.kind == VmValue::EMPTY) throw Error("vm: read of an empty register");
    VmValue out;
    if (in.op == VM_NEG) {
      if (a.kind == VmValue::PUBLIC) { out.kind = VmValue::PUBLIC; out.pub = pub_op(COCG_OP_NEG, a.pub, nullptr); }
      else {
        out.kind = VmValue::SHARED;
        out.sh = driver.alloc_share(batch);
        for (int k = 0; k < K; k++) vop(COCG_OP_NEG, comp(a.sh, k), nullptr, comp(out.sh, k));
      }
    } else {
      const VmValue& b = at(in.rhs);
      if (b.kind == VmValue::EMPTY) throw Error("vm: read of an empty register");
      const bool ap = a.kind == VmValue::PUBLIC, bp = b.kind == VmValue::PUBLIC;
      switch (in.op) {
        case VM_ADD:
          if (ap && bp) { out.kind = VmValue::PUBLIC; out.pub = pub_op(COCG_OP_ADD, a.pub, &b.pub); }
          else if (ap) { out.kind = VmValue::SHARED; out.sh = add_public(b.sh, a.pub, false); }
          else if (bp) { out.kind = VmValue::SHARED; out.sh = add_public(a.sh, b.pub, false); }
          else {
            out.kind = VmValue::SHARED;
            out.sh = driver.alloc_share(batch);
            for (int k = 0; k < K; k++) vop(COCG_OP_ADD, comp(a.sh, k), &comp(b.sh, k), comp(out.sh, k));
          }
          break;
        case VM_SUB:
          if (ap && bp) { out.kind = VmValue::PUBLIC; out.pub = pub_op(COCG_OP_SUB, a.pub, &b.pub); }
          else if (ap) { out.kind = VmValue::SHARED; out.sh = add_public(b.sh, a.pub, true); }  // a + (-b)
          else if (bp) {
            DevVec nb = pub_op(COCG_OP_NEG, b.pub, nullptr);
            out.kind = VmValue::SHARED;
            out.sh = add_public(a.sh, nb, false);
            driver.release(nb);
          } else {
            out.kind = VmValue::SHARED;
            out.sh = driver.alloc_share(batch);
            for (int k = 0; k < K; k++) vop(COCG_OP_SUB, comp(a.sh, k), &comp(b.sh, k), comp(out.sh, k));
          }
          break;
        case VM_MUL:
          if (ap && bp) { out.kind = VmValue::PUBLIC; out.pub = pub_op(COCG_OP_MUL, a.pub, &b.pub); }
          else if (ap) { out.kind = VmValue::SHARED; out.sh = mul_public(b.sh, a.pub); }
          else if (bp) { out.kind = VmValue::SHARED; out.sh = mul_public(a.sh, b.pub); }
          else { out.kind = VmValue::SHARED; out.sh = mul_shared(a.sh, b.sh); }
          break;
        case VM_DIV:  // a * b^-1 (witness_extension_impl.rs:170-200); a zero divisor raises the reference's error
          if (bp) {
            DevVec inv = driver.alloc(batch);
            check(driver.ctx, cocg_d2d(driver.ctx, inv.p, b.pub.p, batch * 32), "cocg_d2d");
            size_t zeros = 0;
            check(driver.ctx, cocg_vec_inv(driver.ctx, inv.p, inv.p, batch, &zeros), "cocg_vec_inv");
            if (zeros) { driver.release(inv); throw Error("Cannot invert zero"); }
            if (ap) { out.kind = VmValue::PUBLIC; out.pub = pub_op(COCG_OP_MUL, a.pub, &inv); }
            else { out.kind = VmValue::SHARED; out.sh = mul_public(a.sh, inv); }
            driver.release(inv);
          } else {
            FieldShareVec binv = inv_shared(b.sh);
            out.kind = VmValue::SHARED;
            if (ap) out.sh = mul_public(binv, a.pub);
            else out.sh = mul_shared(a.sh, binv);
            driver.release(binv);
          }
          break;
        default:
          throw Error("vm: unknown opcode");
      }
    }
    VmValue& dst = at(in.dst);
    clear(dst);
    dst = out;
  }
};

}  // namespace cohost
