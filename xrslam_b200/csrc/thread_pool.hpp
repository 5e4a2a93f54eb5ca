// Minimal persistent thread pool: parallel_for over independent sessions (host bookkeeping only).
// Every parallel_for publishes a job object of its own (function, item count, next-item and pending counters);
// workers pick the current job up under the mutex and touch nothing but that job, and parallel_for returns only
// when every item is done AND every worker that picked the job up has let go of it, so a straggler can never
// see the fields of the next job half-written (ThreadSanitizer-clean, tests/cuemu replay build).
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
        Job job;
        job.fn = &fn; job.n = n; job.pending = n;
        {
            std::unique_lock<std::mutex> l(m_);
            job_ = &job;
            generation_++;
        }
        cv_.notify_all();
        run_some(job);
        std::unique_lock<std::mutex> l(m_);
        job_ = nullptr;                                   // no new worker can pick it up from here on
        done_cv_.wait(l, [&] { return job.pending == 0 && job.users == 0; });
    }

  private:
    struct Job {
        const std::function<void(int)> *fn = nullptr;
        int n = 0;
        std::atomic<int> next{0};
        int pending = 0;       // items not finished yet   (guarded by m_)
        int users = 0;         // workers inside run_some  (guarded by m_)
    };
    void run_some(Job &job) {
        while (true) {
            const int i = job.next.fetch_add(1);
            if (i >= job.n) break;
            (*job.fn)(i);
            std::unique_lock<std::mutex> l(m_);
            if (--job.pending == 0) done_cv_.notify_all();
        }
    }
    void worker() {
        unsigned long long seen = 0;
        while (true) {
            Job *job;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return stop_ || (generation_ != seen && job_ != nullptr); });
                if (stop_) return;
                seen = generation_;
                job = job_;
                job->users++;
            }
            run_some(*job);
            {
                std::unique_lock<std::mutex> l(m_);
                if (--job->users == 0) done_cv_.notify_all();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    Job *job_ = nullptr;
    unsigned long long generation_ = 0;
    bool stop_ = false;
};

}  // namespace xvs
