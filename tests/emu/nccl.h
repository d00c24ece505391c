// tests/emu/nccl.h -- TEST INFRASTRUCTURE ONLY: the handful of NCCL declarations bcalm_b200/csrc/nccl_dyn.h needs to compile
// in the CPU emulation build (tests/emu/cuda_runtime.h).  NCCL itself is bound with dlopen at run time, and the emulation
// build never initialises a communicator: sharded mode is out of its reach.
#pragma once
#include <stddef.h>
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1, ncclSystemError = 2, ncclInternalError = 3 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclChar = 0, ncclUint8 = 1, ncclInt32 = 2, ncclInt = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5 } ncclDataType_t;
