// seaweedfs_b200/csrc/io_pool.h — a small blocking fork-join pool for host-side work that the GPU cannot do:
// the k preads / k+m pwrites of a stripe in the file pipeline (ec_files.cc) and the bounce copies between
// pageable caller memory and the pinned staging ring at the Encoder seam (engine.cc apply_host).  No GF arithmetic
// ever runs here.  tests/test_iopool.py compiles this class on its own under ThreadSanitizer and AddressSanitizer.
#pragma once
#include <algorithm>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace swec {

// A few I/O threads shared by the reader and the writer side: every shard file is independent, so
// the k preads of a stripe (and the k+m pwrites of a finished one) run concurrently.  One thread
// doing them serially tops out near 1.5 GB/s even on RAM-backed files.
class IoPool {
  public:
    explicit IoPool(size_t n) {
        for (size_t i = 0; i < n; i++) threads_.emplace_back([this] { loop(); });
    }
    ~IoPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    // run fn(0..n-1) across the pool (the caller takes a share too); returns the first non-zero result
    int parallel_for(int n, const std::function<int(int)>& fn) {
        if (n <= 0) return 0;
        Batch b;  // lives on this stack frame: nobody may touch it once finished == n has been observed
        b.fn = &fn;
        b.n = n;
        std::unique_lock<std::mutex> lk(mu_);
        batches_.push_back(&b);
        cv_.notify_all();
        work(&b, lk);
        b.done_cv.wait(lk, [&] { return b.finished == b.n; });
        batches_.erase(std::find(batches_.begin(), batches_.end(), &b));
        return b.rc;
    }

  private:
    struct Batch {
        const std::function<int(int)>* fn = nullptr;
        int n = 0, next = 0, finished = 0, rc = 0;
        std::condition_variable done_cv;
    };
    // Called and returns with mu_ held.  A batch is only dereferenced while the lock has been held
    // continuously since we last saw it unfinished (its owner cannot return without the lock).
    void work(Batch* b, std::unique_lock<std::mutex>& lk) {
        while (b->next < b->n) {
            const int i = b->next++;
            const std::function<int(int)>* fn = b->fn;
            lk.unlock();
            const int rc = (*fn)(i);
            lk.lock();
            if (rc && !b->rc) b->rc = rc;
            if (++b->finished == b->n) {
                b->done_cv.notify_all();
                return;  // the owner may destroy the batch as soon as we release the lock
            }
        }
    }
    void loop() {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            Batch* b = nullptr;
            cv_.wait(lk, [&] {
                if (stop_) return true;
                for (Batch* x : batches_)
                    if (x->next < x->n) { b = x; return true; }
                return false;
            });
            if (stop_) return;
            if (b) work(b, lk);
        }
    }
    std::vector<std::thread> threads_;
    std::deque<Batch*> batches_;
    std::mutex mu_;
    std::condition_variable cv_;
    bool stop_ = false;
};

// ---- end of IoPool

}  // namespace swec
