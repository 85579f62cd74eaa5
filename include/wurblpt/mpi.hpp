/* mpi.hpp -- MPICoordinator (mirror of the public interface of reference libwurblpt/mpi.hpp:152-289).
 *
 * The reference distributes 4096-pixel blocks over MPI ranks: a coordinator thread on rank 0 answers block requests
 * (mpi.hpp:74-107), the ranks render their blocks with OpenMP and send the pixels back (mpi.hpp:133-149). Here ONE
 * process drives all GPUs of the box: Scene::updateBVH() replicates the scene on every device (WURBLPT_DEVICES
 * selects them, default: all), mcpt() renders one frame on all of them, the blocks of work items come from one atomic
 * counter in the first GPU's memory, and the frames are combined device to device (wurblpt_b200.h, WPT_DEVICE_ALL).
 * worldSize() is therefore the number of GPUs, there is one (main) process, and the block interface below has
 * nothing left to hand out: it describes the whole image as one block, like the reference built without MPI
 * (mpi.hpp:241-254).
 *
 * Starting this program under mpirun / srun with several ranks would make every rank render the whole frame on the
 * same GPUs and write the same output files; the constructor refuses to do that.
 */
#pragma once

#include <cstdio>
#include <cstdlib>

#include "../wurblpt_b200.h"

namespace WurblPT {

class MPICoordinator
{
private:
    unsigned int _blockSize;
    char _processId[32];
    unsigned int _width, _height;
    bool _blockHandedOut;

    static int launcherWorldSize()
    {
        static const char* names[] = { "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "MV2_COMM_WORLD_SIZE", "PMIX_SIZE" };
        for (const char* name : names) {
            const char* v = std::getenv(name);
            if (v && std::atoi(v) > 1)
                return std::atoi(v);
        }
        return 1;
    }

public:
    MPICoordinator(unsigned int blockSize = 4096) : _blockSize(blockSize), _width(0), _height(0), _blockHandedOut(false)
    {
        std::snprintf(_processId, sizeof(_processId), "process 0");
        if (launcherWorldSize() > 1 && !std::getenv("WURBLPT_ALLOW_MULTI_PROCESS")) {
            std::fprintf(stderr, "WurblPT (B200 build): started as one of %d MPI ranks. This build renders on all GPUs of "
                    "the box from ONE process (set WURBLPT_DEVICES to choose them); start it without mpirun.\n",
                    launcherWorldSize());
            std::abort();
        }
    }

    unsigned int blockSize() const { return _blockSize; }
    bool mainProcess() const { return true; }
    /* the reference: number of MPI ranks (mpi.hpp:286); here: number of GPUs that render a frame */
    int worldSize() const
    {
        const char* single = std::getenv("WURBLPT_DEVICE");    /* one device chosen explicitly (scene.hpp) */
        const int n = (single && *single ? 1 : wpt_planned_device_count());
        return n < 1 ? 1 : n;
    }
    const char* processId() const { return _processId; }

    /* the reference's block interface (mpi.hpp:220-278) in its single-process form */
    void init(unsigned int width, unsigned int height, float* /* pixelData */, unsigned int /* pixelComponents */)
    {
        _width = width;
        _height = height;
        _blockHandedOut = false;
    }
    void getBlock(unsigned int* blockStart, unsigned int* blockSize)
    {
        *blockStart = 0;
        *blockSize = (_blockHandedOut ? 0u : _width * _height);
        _blockHandedOut = true;
    }
    void submitBlock(unsigned int /* blockStart */, unsigned int /* blockSize */) {}
    void finish() {}
};

}
