// Minimal persistent thread pool: parallel_for over independent sessions (host bookkeeping only).
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace xvs {

class ThreadPool {
  public:
    explicit ThreadPool(int n_threads) {
        if (n_threads < 1) n_threads = 1;
        for (int i = 0; i + 1 < n_threads; ++i) workers_.emplace_back([this] { worker(); });
    }
    ~ThreadPool() {
        {
            std::unique_lock<std::mutex> l(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    int size() const { return (int)workers_.size() + 1; }

    // Calls fn(i) for i in [0, n); returns when all calls have finished.  The caller participates.
    void parallel_for(int n, const std::function<void(int)> &fn) {
        if (n <= 0) return;
        if (workers_.empty() || n == 1) { for (int i = 0; i < n; ++i) fn(i); return; }
        {
            std::unique_lock<std::mutex> l(m_);
            fn_ = &fn; n_ = n; next_.store(0); pending_ = n; generation_++;
        }
        cv_.notify_all();
        run_some();
        std::unique_lock<std::mutex> l(m_);
        done_cv_.wait(l, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

  private:
    void run_some() {
        while (true) {
            const int i = next_.fetch_add(1);
            if (i >= n_) break;
            (*fn_)(i);
            std::unique_lock<std::mutex> l(m_);
            if (--pending_ == 0) done_cv_.notify_all();
        }
    }
    void worker() {
        unsigned long long seen = 0;
        while (true) {
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return stop_ || (generation_ != seen && fn_ != nullptr); });
                if (stop_) return;
                seen = generation_;
            }
            run_some();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)> *fn_ = nullptr;
    int n_ = 0, pending_ = 0;
    std::atomic<int> next_{0};
    unsigned long long generation_ = 0;
    bool stop_ = false;
};

}  // namespace xvs
