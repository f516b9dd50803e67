// host_copier.hpp -- multi-threaded memcpy between the caller's pageable arrays and the engine's pinned staging buffers.
// A pageable cudaMemcpyAsync is staged by the driver on one thread (measured 11 GB/s on the B200 box: 39 ms for the 440 MB
// of a 1e7-state DoubleIntegrator_implicit_tb batch); splitting the same copies over a few host threads and issuing the
// DMA from pinned memory brings the link back as the limit.  Not on the device path; nothing here touches CUDA.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace {

class HostCopier {
public:
	struct Piece {
		void *dst;
		const void *src;
		size_t bytes;
	};

	explicit HostCopier(int threads)
	{
		for (int i = 1; i < threads; i++) workers_.emplace_back([this]() { run(); });
	}
	~HostCopier()
	{
		{
			std::lock_guard<std::mutex> lk(m_);
			stop_ = true;
		}
		cv_.notify_all();
		for (std::thread &t : workers_) t.join();
	}
	HostCopier(const HostCopier &) = delete;
	HostCopier &operator=(const HostCopier &) = delete;

	static int default_threads()
	{
		if (const char *v = getenv("ASIF_B200_COPY_THREADS")) {
			const int t = atoi(v);
			if (t >= 1) return t < 64 ? t : 64;
		}
		const unsigned hw = std::thread::hardware_concurrency();
		const int t = hw ? (int)hw : 4;
		return t < 8 ? t : 8;
	}

	// copies every job, split into 1 MiB pieces over the workers and the calling thread; returns when all are done
	void copy(const Piece *jobs, int n_jobs)
	{
		constexpr size_t PIECE = (size_t)1 << 20;
		std::vector<Piece> pcs;
		for (int j = 0; j < n_jobs; j++)
			for (size_t o = 0; o < jobs[j].bytes; o += PIECE)
				pcs.push_back({(char *)jobs[j].dst + o, (const char *)jobs[j].src + o, jobs[j].bytes - o < PIECE ? jobs[j].bytes - o : PIECE});
		if (workers_.empty() || pcs.size() < 2) {
			for (const Piece &p : pcs) memcpy(p.dst, p.src, p.bytes);
			return;
		}
		{
			std::unique_lock<std::mutex> lk(m_);
			cv_done_.wait(lk, [this]() { return active_ == 0; }); // no worker is still looking at the previous piece list
			pieces_.swap(pcs);
			next_.store(0);
			done_.store(0);
			total_ = pieces_.size();
			gen_++;
		}
		cv_.notify_all();
		work();
		std::unique_lock<std::mutex> lk(m_);
		cv_done_.wait(lk, [this]() { return done_.load() == total_; });
	}

private:
	void work()
	{
		for (;;) {
			const size_t i = next_.fetch_add(1);
			if (i >= total_) break;
			const Piece &p = pieces_[i];
			memcpy(p.dst, p.src, p.bytes);
			if (done_.fetch_add(1) + 1 == total_) {
				std::lock_guard<std::mutex> lk(m_);
				cv_done_.notify_all();
			}
		}
	}
	void run()
	{
		unsigned long long seen = 0;
		for (;;) {
			{
				std::unique_lock<std::mutex> lk(m_);
				cv_.wait(lk, [&]() { return stop_ || gen_ != seen; });
				if (stop_) return;
				seen = gen_;
				active_++; // total_ and pieces_ are stable until active_ is back to 0
			}
			work();
			{
				std::lock_guard<std::mutex> lk(m_);
				active_--;
			}
			cv_done_.notify_all();
		}
	}

	std::vector<std::thread> workers_;
	std::mutex m_;
	std::condition_variable cv_, cv_done_;
	std::vector<Piece> pieces_;
	std::atomic<size_t> next_{0}, done_{0};
	size_t total_ = 0; // written under m_ while active_ == 0 and the caller is outside work()
	unsigned long long gen_ = 0;
	int active_ = 0;
	bool stop_ = false;
};

} // namespace
