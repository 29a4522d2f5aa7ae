// The REP3 party network the drivers talk to, and its in-process implementation.
//
// Mirrors the reference's `Rep3Network` trait (/root/reference/mpc-core/src/protocols/rep3/network.rs:13-64:
// get_id, send/recv, send_next(_many), recv_prev(_many)) and the test double the reference runs its provers on,
// `Rep3TestNetwork` / `PartyTestNetwork` (/root/reference/tests/src/rep3_network.rs:6-166: one unbounded byte channel
// per ordered pair of parties).  The MPC rounds stay on the host (north_star): a message is a host byte buffer.  Large
// share vectors travel in page-locked buffers so the sender's D2H and the receiver's H2D run at PCIe rate; the buffer
// is handed over by reference count instead of being copied a second time.
// Errors: a closed channel surfaces as cohost::Error (the reference's io::ErrorKind::BrokenPipe, network.rs:157-159).
#pragma once
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <vector>

#include "types.hpp"

namespace cohost {

struct Message {
  std::shared_ptr<void> data;  // host memory (pinned for share vectors)
  size_t bytes = 0;
  void* device = nullptr;      // device-exchange mode: an HBM buffer whose ownership moves to the receiver (data is null)
};

class Channel {  // unbounded MPSC byte channel (std::sync::mpsc in the reference's test network)
 public:
  void send(Message m) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (closed_) throw Error("network: send on a closed channel (BrokenPipe)");
      q_.push_back(std::move(m));
    }
    cv_.notify_one();
  }
  Message recv() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return !q_.empty() || closed_; });
    if (q_.empty()) throw Error("network: peer hung up (BrokenPipe)");
    Message m = std::move(q_.front());
    q_.pop_front();
    return m;
  }
  void close() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      closed_ = true;
    }
    cv_.notify_all();
  }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<Message> q_;
  bool closed_ = false;
};

class Rep3Network {
 public:
  virtual ~Rep3Network() {}
  virtual int get_id() const = 0;
  virtual void send(int target, Message m) = 0;
  virtual Message recv(int from) = 0;
  // true when all parties of this network live in one process on ONE GPU and the caller opted in: share vectors may then be
  // handed over as HBM buffers instead of being staged through pinned host memory (the default, which is what a party that has
  // to reach a NIC does)
  virtual bool device_exchange() const { return false; }
  void send_next(Message m) { send((get_id() + 1) % 3, std::move(m)); }
  Message recv_prev() { return recv((get_id() + 2) % 3); }
  // small values: copied into a heap buffer
  void send_next_bytes(const void* p, size_t n) {
    std::shared_ptr<void> buf(new uint8_t[n ? n : 1], [](void* q) { delete[] (uint8_t*)q; });
    memcpy(buf.get(), p, n);
    send_next(Message{buf, n});
  }
  void recv_prev_bytes(void* p, size_t n) {
    Message m = recv_prev();
    if (m.bytes != n) throw Error("network: invalid number of bytes received");  // rep3.rs:663-668
    memcpy(p, m.data.get(), n);
  }
};

// Cross-GPU leg of the REP3 network (block mode, one process per GPU): the witness maps of the three parties run on different ranks,
// so their mul_vec payloads travel GPU to GPU (NCCL send / recv over NVLink, issued by the caller's callback) instead of through the
// in-process channels.  Party threads of one rank post their transfers of a round; the last one to arrive hands the whole batch to the
// callback (one grouped NCCL call per rank and round, so sends and receives cannot deadlock) and releases the others.
struct CommOp {
  int dir;        // 0 = send, 1 = receive
  int peer;       // rank
  void* dptr;     // DEVICE buffer on this rank's GPU
  size_t bytes;
};
typedef int (*CommCallback)(void* user, const CommOp* ops, int nops);
class DeviceBridge {
 public:
  DeviceBridge(int rank, const int owner_of_party[3], CommCallback cb, void* user) : rank_(rank), cb_(cb), user_(user) {
    for (int p = 0; p < 3; p++) owner_[p] = owner_of_party[p];
    for (int p = 0; p < 3; p++)  // a party of this rank posts iff one of its ring neighbours lives elsewhere
      if (owner_[p] == rank_ && (owner_[(p + 1) % 3] != rank_ || owner_[(p + 2) % 3] != rank_)) posters_++;
  }
  int rank() const { return rank_; }
  int owner(int party) const { return owner_[party]; }
  bool local(int party) const { return owner_[party] == rank_; }
  void post(const CommOp& op) {
    std::lock_guard<std::mutex> lk(mu_);
    ops_.push_back(op);
  }
  // called once per round by every posting party thread
  void flush() {
    std::unique_lock<std::mutex> lk(mu_);
    if (failed_) throw Error("device bridge: an earlier transfer failed");
    const uint64_t gen = gen_;
    if (++arrived_ == posters_) {
      int rc = cb_(user_, ops_.data(), (int)ops_.size());
      ops_.clear();
      arrived_ = 0;
      gen_++;
      if (rc) failed_ = true;
      cv_.notify_all();
      if (rc) throw Error("device bridge: the transfer callback failed");
    } else {
      cv_.wait(lk, [&] { return gen_ != gen || failed_; });
      if (failed_) throw Error("device bridge: a transfer failed");
    }
  }
  void abort() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      failed_ = true;
    }
    cv_.notify_all();
  }

 private:
  int rank_, owner_[3], posters_ = 0, arrived_ = 0;
  CommCallback cb_;
  void* user_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::vector<CommOp> ops_;
  uint64_t gen_ = 0;
  bool failed_ = false;
};

class Rep3TestNetwork;
class PartyTestNetwork : public Rep3Network {
 public:
  PartyTestNetwork(Rep3TestNetwork* net, int id) : net_(net), id_(id) {}
  int get_id() const override { return id_; }
  void send(int target, Message m) override;
  Message recv(int from) override;
  bool device_exchange() const override;

 private:
  Rep3TestNetwork* net_;
  int id_;
};

class Rep3TestNetwork {
 public:
  Rep3TestNetwork() {
    for (int i = 0; i < 3; i++) parties_[i].reset(new PartyTestNetwork(this, i));
  }
  PartyTestNetwork* party(int i) { return parties_[i].get(); }
  Channel& chan(int from, int to) { return ch_[from][to]; }
  bool device_exchange = false;
  void close_all() {
    for (auto& row : ch_)
      for (auto& c : row) c.close();
  }

 private:
  Channel ch_[3][3];
  std::unique_ptr<PartyTestNetwork> parties_[3];
};

inline void PartyTestNetwork::send(int target, Message m) { net_->chan(id_, target).send(std::move(m)); }
inline Message PartyTestNetwork::recv(int from) { return net_->chan(from, id_).recv(); }
inline bool PartyTestNetwork::device_exchange() const { return net_->device_exchange; }

}  // namespace cohost
