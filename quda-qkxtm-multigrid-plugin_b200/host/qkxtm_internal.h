// qkxtm_internal.h -- what the translation units of libqkxtm_tmq.so share besides the public headers
#pragma once
namespace quda {
const int *qkxtm_local_extent();        // GK_localL after init_qudaQKXTM
long long qkxtm_local_volume();         // GK_localVolume
void qkxtm_raise(const char *msg);      // errorQuda from outside qudaQKXTM_tmq.cpp: prints and aborts (or calls the installed handler)
}
