// Native batch prefetcher: a C++ worker thread that, per mini-batch, gathers the
// sampled uint8 images from the decoded cache into a pinned staging ring, draws the
// RandomResizedCrop / RandomRotation / RandomHorizontalFlip parameters (torchvision's
// distributions) and publishes the slot.  It never takes the Python GIL, so it cannot stall the
// thread that launches the training step (a Python producer thread does: with the GIL's 5 ms switch
// interval the launch loop of a 7.6 ms step was held up and end-to-end throughput halved).
//
// Slot life cycle:  free -> (worker fills) -> ready -> (consumer: next()) -> in use ->
// (consumer: release(slot, cuda_event)) -> free; the worker cudaEventSynchronize()s the event the
// consumer recorded after its host->device copy before it overwrites the slot.
#include <cuda_runtime.h>

#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

namespace b200 {

struct PrefetchConfig {
  const uint8_t* cache;      // [N][H][W][3]
  const int64_t* labels;     // [N]
  int H, W, mb, slots;
  uint8_t* ring_img;         // [slots][mb][H][W][3]   (pinned)
  float* ring_par;           // [slots][mb][8]
  int64_t* ring_lab;         // [slots][mb]
  bool train;
  float scale_lo, scale_hi, ratio_lo, ratio_hi, degrees;
};

class BatchPrefetcher {
 public:
  explicit BatchPrefetcher(const PrefetchConfig& c) : cfg_(c), events_(c.slots, nullptr) {
    for (int s = 0; s < c.slots; ++s) free_.push_back(s);
  }
  ~BatchPrefetcher() { stop(); }

  void start_epoch(std::vector<int64_t> indices, uint64_t seed) {
    stop();
    {
      std::lock_guard<std::mutex> lk(mu_);
      indices_ = std::move(indices);
      ready_.clear();
      done_ = false;
      quit_ = false;
      free_.clear();
      for (int s = 0; s < cfg_.slots; ++s) free_.push_back(s);
    }
    rng_.seed(seed);
    worker_ = std::thread([this] { run(); });
  }

  // Blocks until a batch is ready.  Returns false at the end of the epoch.
  bool next(int& slot, int& count) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_ready_.wait(lk, [this] { return !ready_.empty() || done_; });
    if (ready_.empty()) return false;
    slot = ready_.front().first;
    count = ready_.front().second;
    ready_.pop_front();
    return true;
  }

  void release(int slot, cudaEvent_t ev) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      events_[slot] = ev;
      free_.push_back(slot);
    }
    cv_free_.notify_one();
  }

  void stop() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      quit_ = true;
    }
    cv_free_.notify_all();
    if (worker_.joinable()) worker_.join();
  }

 private:
  double uni() { return std::generate_canonical<double, 53>(rng_); }

  void sample_params(float* p) {
    const PrefetchConfig& c = cfg_;
    if (!c.train) {
      p[0] = 0; p[1] = 0; p[2] = (float)c.H; p[3] = (float)c.W; p[4] = 1.f; p[5] = 0.f; p[6] = 0.f; p[7] = 0.f;
      return;
    }
    const double area = (double)c.H * c.W, llo = std::log(c.ratio_lo), lhi = std::log(c.ratio_hi);
    double cw = 0, ch = 0;
    bool ok = false;
    for (int t = 0; t < 10; ++t) {                      // every draw consumed: fixed RNG stride
      const double ta = area * (c.scale_lo + (c.scale_hi - c.scale_lo) * uni());
      const double ar = std::exp(llo + (lhi - llo) * uni());
      const double w = std::nearbyint(std::sqrt(ta * ar)), h = std::nearbyint(std::sqrt(ta / ar));
      if (!ok && w > 0 && w <= c.W && h > 0 && h <= c.H) { cw = w; ch = h; ok = true; }
    }
    const double u0 = uni(), u1 = uni(), ua = uni(), uf = uni();
    double top, left;
    if (ok) {
      top = std::fmin(std::floor(u0 * (c.H - ch + 1)), c.H - ch);
      left = std::fmin(std::floor(u1 * (c.W - cw + 1)), c.W - cw);
    } else {                                            // torchvision's centre-crop fallback
      const double in_ratio = (double)c.W / c.H;
      if (in_ratio < c.ratio_lo) { cw = c.W; ch = std::nearbyint(cw / c.ratio_lo); }
      else if (in_ratio > c.ratio_hi) { ch = c.H; cw = std::nearbyint(ch * c.ratio_hi); }
      else { cw = c.W; ch = c.H; }
      top = std::floor((c.H - ch) / 2); left = std::floor((c.W - cw) / 2);
    }
    const double theta = (-c.degrees + 2.0 * c.degrees * ua) * 3.14159265358979323846 / 180.0;
    p[0] = (float)top; p[1] = (float)left; p[2] = (float)ch; p[3] = (float)cw;
    p[4] = (float)std::cos(theta); p[5] = (float)std::sin(theta); p[6] = uf < 0.5 ? 1.f : 0.f; p[7] = 0.f;
  }

  void run() {
    const PrefetchConfig& c = cfg_;
    const size_t img_bytes = (size_t)c.H * c.W * 3;
    const size_t n = indices_.size();
    for (size_t b0 = 0; b0 < n; b0 += c.mb) {
      int slot;
      cudaEvent_t ev;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_free_.wait(lk, [this] { return !free_.empty() || quit_; });
        if (quit_) {                      // stopped mid-epoch: a consumer blocked in next() must wake up
          done_ = true;
          lk.unlock();
          cv_ready_.notify_all();
          return;
        }
        slot = free_.front();
        free_.pop_front();
        ev = events_[slot];
        events_[slot] = nullptr;
      }
      if (ev) cudaEventSynchronize(ev);               // previous occupant's H2D copy has finished
      const int k = (int)std::min<size_t>(c.mb, n - b0);
      uint8_t* img = c.ring_img + (size_t)slot * c.mb * img_bytes;
      float* par = c.ring_par + (size_t)slot * c.mb * 8;
      int64_t* lab = c.ring_lab + (size_t)slot * c.mb;
      for (int i = 0; i < k; ++i) {
        const int64_t src = indices_[b0 + i];
        std::memcpy(img + (size_t)i * img_bytes, c.cache + (size_t)src * img_bytes, img_bytes);
        lab[i] = c.labels[src];
        sample_params(par + (size_t)i * 8);
      }
      {
        std::lock_guard<std::mutex> lk(mu_);
        ready_.emplace_back(slot, k);
      }
      cv_ready_.notify_one();
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      done_ = true;
    }
    cv_ready_.notify_all();
  }

  PrefetchConfig cfg_;
  std::vector<int64_t> indices_;
  std::deque<int> free_;
  std::deque<std::pair<int, int>> ready_;
  std::vector<cudaEvent_t> events_;
  std::mutex mu_;
  std::condition_variable cv_free_, cv_ready_;
  std::thread worker_;
  std::mt19937_64 rng_;
  bool done_ = false, quit_ = false;
};

BatchPrefetcher* prefetcher_create(const PrefetchConfig& c) { return new BatchPrefetcher(c); }
void prefetcher_destroy(BatchPrefetcher* p) { delete p; }
void prefetcher_start(BatchPrefetcher* p, std::vector<int64_t> idx, uint64_t seed) { p->start_epoch(std::move(idx), seed); }
bool prefetcher_next(BatchPrefetcher* p, int& slot, int& count) { return p->next(slot, count); }
void prefetcher_release(BatchPrefetcher* p, int slot, uint64_t cuda_event) {
  p->release(slot, reinterpret_cast<cudaEvent_t>(cuda_event));
}
void prefetcher_stop(BatchPrefetcher* p) { p->stop(); }

}  // namespace b200
