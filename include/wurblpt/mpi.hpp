/* mpi.hpp -- MPICoordinator (mirror of the public interface of reference libwurblpt/mpi.hpp:152-289).
 *
 * The reference distributes 4096-pixel blocks over MPI ranks. On B200 the image is split into
 * small pixel tiles that warps pull from a device-side atomic ticket counter, and across GPUs the
 * tiles are interleaved (one process per GPU, frames combined with one NCCL reduce; see
 * wurblpt_b200/multi.py). Inside one process this class therefore has nothing to coordinate; it keeps
 * the reference's interface so that applications written for MPI compile and run unchanged on one GPU.
 */
#pragma once

#include <cstdio>

namespace WurblPT {

class MPICoordinator
{
private:
    unsigned int _blockSize;
    char _processId[32];

public:
    MPICoordinator(unsigned int blockSize = 4096) : _blockSize(blockSize)
    {
        std::snprintf(_processId, sizeof(_processId), "process 0");
    }

    unsigned int blockSize() const { return _blockSize; }
    bool mainProcess() const { return true; }
    int worldSize() const { return 1; }
    const char* processId() const { return _processId; }
};

}
