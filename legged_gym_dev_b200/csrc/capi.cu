// Library-level entry points of the C ABI: version, error string, struct sizes.
#include "common.cuh"
#include "../../include/b200gym.h"
#include <string.h>

static thread_local char g_err[512] = "";

void b200gym_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {

int b200gym_version(void) { return B200GYM_VERSION; }

const char* b200gym_last_error(void) { return g_err; }

int b200gym_sizeof(const char* name) {
    if (!name) return -1;
    if (!strcmp(name, "B200LeggedParams")) return (int)sizeof(B200LeggedParams);
    if (!strcmp(name, "B200LeggedBuffers")) return (int)sizeof(B200LeggedBuffers);
    if (!strcmp(name, "B200MlpParams")) return (int)sizeof(B200MlpParams);
    if (!strcmp(name, "B200PpoLossParams")) return (int)sizeof(B200PpoLossParams);
    if (!strcmp(name, "B200RomParams")) return (int)sizeof(B200RomParams);
    if (!strcmp(name, "B200RomState")) return (int)sizeof(B200RomState);
    if (!strcmp(name, "B200RomFamilyParams")) return (int)sizeof(B200RomFamilyParams);
    if (!strcmp(name, "B200HopperTorqueParams")) return (int)sizeof(B200HopperTorqueParams);
    if (!strcmp(name, "B200HopperTorqueBuffers")) return (int)sizeof(B200HopperTorqueBuffers);
    if (!strcmp(name, "B200HopperEnvParams")) return (int)sizeof(B200HopperEnvParams);
    if (!strcmp(name, "B200HopperEnvBuffers")) return (int)sizeof(B200HopperEnvBuffers);
    if (!strcmp(name, "B200HopperObsParams")) return (int)sizeof(B200HopperObsParams);
    if (!strcmp(name, "B200PeerPtrs")) return (int)sizeof(B200PeerPtrs);
    if (!strcmp(name, "B200PeerBases")) return (int)sizeof(B200PeerBases);
    if (!strcmp(name, "B200GemmProblem")) return (int)sizeof(B200GemmProblem);
    if (!strcmp(name, "B200PackTable")) return (int)sizeof(B200PackTable);
    if (!strcmp(name, "B200ChainNet")) return (int)sizeof(B200ChainNet);
    if (!strcmp(name, "B200OptParams")) return (int)sizeof(B200OptParams);
    return -1;
}

}  // extern "C"
