// Instantiation of the per-group host routines and kernels for Mnt6G1 (see group_ops.cuh).
#include "group_ops.cuh"

extern const GroupOps b200msm_ops_mnt6g1;
const GroupOps b200msm_ops_mnt6g1 = make_group_ops<mnt753::Mnt6G1>();
