// nccl.h — HOST STAND-IN (tests/emul only): the slice of the NCCL API that cajitafluids_b200/csrc/halo.cu
// binds with dlsym, implemented in nccl_emul.cpp for several ranks living in ONE process as threads
// (send / recv are buffered copies through in-process mailboxes, collectives go through a barrier).
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

#define NCCL_UNIQUE_ID_BYTES 128
typedef struct
{
    char internal[NCCL_UNIQUE_ID_BYTES];
} ncclUniqueId;
typedef struct cfb_emul_comm* ncclComm_t;
typedef enum { ncclSuccess = 0, ncclInternalError = 3, ncclInvalidArgument = 4 } ncclResult_t;
typedef enum { ncclChar = 0, ncclInt = 2, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;
