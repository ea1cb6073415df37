/* comm_quda.h -- drop-in stand-in for upstream QUDA's communication header: what qkxtm/QKXTM_util.cpp and include/QKXTM_read_conf.h
 * use of it (rank / coordinates of this process in the grid given to initCommsGridQuda). */
#pragma once
#include "quda.h"
struct Topology;                                  /* opaque: the process grid of initCommsGridQuda */
extern Topology *default_topo;
#ifdef __cplusplus
extern "C" {
#endif
const int *comm_coords(const Topology *topo);     /* (x, y, z, t) coordinates of this rank */
const int *comm_dims(const Topology *topo);
void comm_dim_partitioned_set(int dim);
#ifdef __cplusplus
}
#endif
