// TEMPORARY: placeholder symbols until engine.cu / dispatcher.cpp land (same commit series).
#include "../../include/ollamamq_b200.h"
namespace mq { void set_last_error(const char* fmt, ...); }
#define NI(name) mq::set_last_error(#name ": not implemented yet"); return MQ_ERR_NODEV
extern "C" {
int mq_worker_count(void) { return 0; }
int mq_worker_open(int32_t, const mq_model_cfg*, mq_worker**) { NI(mq_worker_open); }
void mq_worker_close(mq_worker*) {}
int mq_worker_load_tensor(mq_worker*, const char*, const void*, size_t) { NI(mq_worker_load_tensor); }
int mq_worker_read_tensor(mq_worker*, const char*, void*, size_t) { NI(mq_worker_read_tensor); }
int mq_worker_init_random(mq_worker*, uint64_t, float) { NI(mq_worker_init_random); }
int mq_worker_capacity(mq_worker*) { return 0; }
int mq_worker_healthy(mq_worker*) { return 0; }
int mq_submit(mq_worker*, const mq_request*, const mq_callbacks*, void*, mq_req**) { NI(mq_submit); }
void mq_cancel(mq_req*) {}
void mq_req_release(mq_req*) {}
int mq_req_get_stats(mq_req*, mq_req_stats*) { NI(mq_req_get_stats); }
int mq_worker_get_stats(mq_worker*, mq_worker_stats*) { NI(mq_worker_get_stats); }
int mq_worker_reset_stats(mq_worker*) { NI(mq_worker_reset_stats); }
int mq_worker_set_timing(mq_worker*, int32_t) { NI(mq_worker_set_timing); }
int mq_debug_forward(mq_worker*, const int32_t*, int32_t, int32_t, float*) { NI(mq_debug_forward); }
int mq_dispatcher_new(mq_worker**, int32_t, int32_t, mq_dispatcher**) { NI(mq_dispatcher_new); }
void mq_dispatcher_free(mq_dispatcher*) {}
int mq_dispatcher_submit(mq_dispatcher*, const char*, const char*, const mq_request*, const mq_callbacks*, void*, uint64_t*) { NI(mq_dispatcher_submit); }
mq_sched* mq_dispatcher_sched(mq_dispatcher*) { return nullptr; }
int mq_dispatcher_set_vip(mq_dispatcher*, const char*) { NI(x); }
int mq_dispatcher_set_boost(mq_dispatcher*, const char*) { NI(x); }
int mq_dispatcher_block_user(mq_dispatcher*, const char*, int32_t) { NI(x); }
int mq_dispatcher_block_ip(mq_dispatcher*, const char*, int32_t) { NI(x); }
int mq_dispatcher_log(mq_dispatcher*, mq_dispatch*, int32_t, int32_t*) { NI(x); }
int mq_dispatcher_drain(mq_dispatcher*, uint32_t) { NI(x); }
}
