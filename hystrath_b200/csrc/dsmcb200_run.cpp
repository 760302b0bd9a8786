// dsmcb200_run -- standalone driver: the time loop of dsmcFoam+
// (applications/solvers/discreteMethods/dsmc/dsmcFoam+/dsmcFoam+.C:92-183) on an unchanged case directory,
// with dsmcCloud::evolve() executed by libdsmcb200 on the GPU.
//
//   dsmcb200_run -case <caseDir> [-parallel] [-device N]
//   dsmcb200_run -initialise -case <caseDir> [-parallel]     the dsmcInitialise+ step: fill the mesh from system/dsmcInitialiseDict
//                                                            and write the start-time cloud (15 significant digits)
//
// -parallel: one process per GPU, rank/size from RANK / WORLD_SIZE / LOCAL_RANK (torchrun-style launchers) or
// OMPI_COMM_WORLD_*; rank k runs processor<k>/ as written by decomposePar.  The ncclUniqueId is handed from
// rank 0 to the others through <caseDir>/.dsmcb200_nccl_id.
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>

#include "dsmc_cloud.h"

static int envInt(const char* a, const char* b, int dflt) {
    const char* v = std::getenv(a);
    if (!v && b) v = std::getenv(b);
    return v ? std::atoi(v) : dflt;
}

int main(int argc, char** argv) {
    std::string caseDir = ".";
    bool parallel = false, dryRun = false, initialise = false;
    int device = -1;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "-case") && i + 1 < argc) caseDir = argv[++i];
        else if (!std::strcmp(argv[i], "-parallel")) parallel = true;
        else if (!std::strcmp(argv[i], "-device") && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "-dryRun")) dryRun = true;
        else if (!std::strcmp(argv[i], "-initialise")) initialise = true;
        else if (!std::strcmp(argv[i], "-renumberCells")) dsmcb200::dsmcCloud::cellOrder(DSMCB200_CELL_ORDER_Z_CURVE);   // renumberMesh in memory: the files keep their labels
        else if (!std::strcmp(argv[i], "-AMR")) { std::fprintf(stderr, "-AMR (dynamic mesh refinement) is outside the scoped path\n"); return 2; }
        else { std::fprintf(stderr, "usage: dsmcb200_run [-initialise] -case <dir> [-parallel] [-device N] [-renumberCells]\n"); return 2; }
    }
    int rank = 0, nRanks = 1;
    if (parallel) {
        rank = envInt("RANK", "OMPI_COMM_WORLD_RANK", 0);
        nRanks = envInt("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", 1);
    }
    if (device < 0) device = parallel ? envInt("LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", rank) : 0;
    try {
        if (dryRun) {  // parse the whole case (dictionaries, mesh, cloud or dsmcInitialiseDict) and report; needs no GPU
            dsmcb200::dsmcCloud probe(caseDir, "dsmc", rank, nRanks, 0, nullptr, true, initialise);
            std::printf("%s", probe.summary().c_str());
            return 0;
        }
        char id[128];
        const void* idPtr = nullptr;
        std::string ncclIdPath;
        if (nRanks > 1) {
            // one file per launch: a run that died before removing its file must not hand a stale id to the next one
            std::string nonce;
            for (const char* v : {"TORCHELASTIC_RUN_ID", "MASTER_PORT", "OMPI_MCA_orte_hnp_uri", "PMI_JOBID", "SLURM_JOB_ID"})
                if (const char* e = std::getenv(v)) { nonce += "_"; for (const char* q = e; *q; ++q) nonce += std::isalnum(static_cast<unsigned char>(*q)) ? *q : '-'; }
            const std::string path = caseDir + "/.dsmcb200_nccl_id" + nonce;
            ncclIdPath = path;
            if (rank == 0) {
                std::remove(path.c_str());
                if (dsmcb200_nccl_unique_id(id) != 0) throw foam::FoamError("cannot create an ncclUniqueId (libnccl.so.2 missing?)");
                FILE* f = std::fopen((path + ".tmp").c_str(), "wb");
                if (!f) throw foam::FoamError("cannot write " + path);
                std::fwrite(id, 1, 128, f);
                std::fclose(f);
                std::rename((path + ".tmp").c_str(), path.c_str());
            } else {
                for (int tries = 0; tries < 600 && !foam::exists(path); ++tries) std::this_thread::sleep_for(std::chrono::milliseconds(100));
                FILE* f = std::fopen(path.c_str(), "rb");
                if (!f || std::fread(id, 1, 128, f) != 128) throw foam::FoamError("cannot read " + path);
                std::fclose(f);
            }
            idPtr = id;
        }
        const bool master = rank == 0;
        if (initialise) {
            // dsmcInitialise+.C:57-88
            dsmcb200::dsmcCloud dsmc(caseDir, "dsmc", rank, nRanks, device, idPtr, false, true);
            if (nRanks > 1 && master) std::remove(ncclIdPath.c_str());
            if (master) std::printf("Initialising dsmc for Time = %s\n\n", dsmc.timeName().c_str());
            dsmc.write();
            if (master) std::printf("\nEnd\n\n");
            return 0;
        }
        if (master) std::printf("\nConstructing dsmcCloud \n");
        dsmcb200::dsmcCloud dsmc(caseDir, "dsmc", rank, nRanks, device, idPtr);
        if (nRanks > 1 && master) std::remove(ncclIdPath.c_str());
        if (master) std::printf("\nStarting time loop\n\n");
        const auto t0 = std::chrono::steady_clock::now();
        const std::clock_t c0 = std::clock();
        int infoCounter = 0;
        long noIteration = 1;
        double lastIter = 0;
        while (dsmc.loop()) {
            infoCounter++;
            const bool talk = infoCounter >= dsmc.nTerminalOutputs();
            if (talk && master) std::printf("Time = %s\n\n", dsmc.timeName().c_str());
            dsmc.evolve();
            if (talk) dsmc.info();
            if (dsmc.outputTime()) dsmc.write();
            const double cpu = double(std::clock() - c0) / CLOCKS_PER_SEC;
            if (talk) {
                const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                if (master)
                    std::printf("\nStage 0.0  ExecutionTime = %.2f s  ClockTime = %d s  Iteration %ld (%g s)\n\n", cpu, int(wall), noIteration,
                                std::max(cpu - lastIter, 1e-3));
                infoCounter = 0;
            }
            lastIter = cpu;
            noIteration += 1;
            std::fflush(stdout);
        }
        if (master) std::printf("End stage 0\n\n");
    } catch (const std::exception& e) {
        // OpenFOAM's FatalError convention: message, then a non-zero exit
        std::fprintf(stderr, "\n\n--> FOAM FATAL ERROR: \n%s\n\nFOAM exiting\n\n", e.what());
        return 1;
    }
    return 0;
}
