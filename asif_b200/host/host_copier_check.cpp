// host_copier_check.cpp -- stress test of csrc/host_copier.hpp (the pageable-array bounce path's thread pool) on the CPU:
// many copy() calls of varying shapes from one caller, results compared byte for byte.  Prints "host_copier ok".
#include "../csrc/host_copier.hpp"
#include <cstdio>
#include <random>

int main()
{
	std::mt19937_64 g(7);
	for (int threads : {1, 2, 5, 8}) {
		HostCopier hc(threads);
		for (int round = 0; round < 300; round++) {
			const int nj = 1 + (int)(g() % 4);
			std::vector<std::vector<unsigned char>> src(nj), dst(nj);
			std::vector<HostCopier::Piece> jobs;
			for (int j = 0; j < nj; j++) {
				const size_t bytes = (round % 7 == 0) ? 0 : (size_t)(g() % (5u << 20)) + (round % 3 == 0 ? 0 : 1);
				src[j].resize(bytes);
				dst[j].assign(bytes, 0);
				for (size_t i = 0; i < bytes; i += 997) src[j][i] = (unsigned char)g();
				if (bytes) src[j][bytes - 1] = (unsigned char)(round + j + 1);
				jobs.push_back({dst[j].data(), src[j].data(), bytes});
			}
			hc.copy(jobs.data(), nj);
			for (int j = 0; j < nj; j++)
				if (src[j] != dst[j]) {
					std::printf("mismatch: threads %d round %d job %d\n", threads, round, j);
					return 1;
				}
		}
	}
	std::printf("host_copier ok\n");
	return 0;
}
