// glibc_rand.hpp — glibc's rand() restated (csrc/host/poisson.cpp): the generator behind the reference's TrueNorth
// threshold jitter (`std::rand() & random_mask`, src/models.cpp:757).
#ifndef SFE_GLIBC_RAND_HPP_
#define SFE_GLIBC_RAND_HPP_

#include <cstdint>

namespace sfe
{
class GlibcRand
{
public:
    explicit GlibcRand(uint32_t seed = 1u); // a process that never calls srand() runs on seed 1
    uint32_t next();                        // the next rand() value, 0 .. 2^31 - 1

private:
    uint32_t state_[31];
    int front_, rear_;
};
} // namespace sfe
#endif
