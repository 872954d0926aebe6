/* fdc_cabi_act.cu -- activity-gated blocks (placeholder until the state machines land): every entry point fails loudly. */
#include "fdc_cabi_internal.h"
using namespace fdc;
#define NI(ret) { fail("not implemented yet"); return ret; }
extern "C" {
fdc_pac* fdc_pac_create(int, float, float, int, float, int, int, int, int, const char*, int, int) NI(0)
int fdc_pac_work_host(fdc_pac*, int, const void*) NI(-1)
int fdc_pac_work_device(fdc_pac*, int, const void*, void*) NI(-1)
int fdc_pac_state(const fdc_pac*, int*, float*) NI(-1)
int fdc_pac_tables(const fdc_pac*, float*) NI(-1)
int fdc_pac_msg_count(const fdc_pac*) NI(-1)
int fdc_pac_msg_get(const fdc_pac*, int, fdc_msg*) NI(-1)
void fdc_pac_msg_clear(fdc_pac*) {}
void fdc_pac_destroy(fdc_pac*) {}
fdc_segdet* fdc_segdet_create(int, int, int, float, float, float, float, float, int, int, int, int, const char*, int, int) NI(0)
int fdc_segdet_work_host(fdc_segdet*, int, const void*) NI(-1)
int fdc_segdet_work_device(fdc_segdet*, int, const void*, void*) NI(-1)
int fdc_segdet_state(const fdc_segdet*, long*, float*) NI(-1)
int fdc_segdet_window(const fdc_segdet*, int, int, float*) NI(-1)
int fdc_segdet_power(const fdc_segdet*, float*) NI(-1)
int fdc_segdet_active(const fdc_segdet*, int, int*) NI(-1)
int fdc_segdet_msg_count(const fdc_segdet*) NI(-1)
int fdc_segdet_msg_get(const fdc_segdet*, int, fdc_msg*) NI(-1)
void fdc_segdet_msg_clear(fdc_segdet*) {}
void fdc_segdet_destroy(fdc_segdet*) {}
fdc_actdet* fdc_actdet_create(int, const float*, int, float, int, int, int, int, const char*, int, float, int, double, int) NI(0)
int fdc_actdet_work_host(fdc_actdet*, int, const void*) NI(-1)
int fdc_actdet_work_device(fdc_actdet*, int, const void*, void*) NI(-1)
int fdc_actdet_nsegments(const fdc_actdet*) NI(-1)
int fdc_actdet_segment(const fdc_actdet*, int, int*) NI(-1)
int fdc_actdet_power(const fdc_actdet*, int, float*) NI(-1)
int fdc_actdet_msg_count(const fdc_actdet*) NI(-1)
int fdc_actdet_msg_get(const fdc_actdet*, int, fdc_msg*) NI(-1)
void fdc_actdet_msg_clear(fdc_actdet*) {}
void fdc_actdet_destroy(fdc_actdet*) {}
}
