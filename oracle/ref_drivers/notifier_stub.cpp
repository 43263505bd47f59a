/*
 * notifier_stub.cpp -- minimal definitions for the Gpu2HostNotifier symbols that the
 * reference's cuda_code.cu links against (gpu2hostnotifier.h:47-52).  The real class is a
 * QObject that re-emits Qt signals; here the static callbacks only count invocations so the
 * harness can check that the reference fired them.  TEST INFRASTRUCTURE ONLY.
 */
#include "gpu2hostnotifier.h"
#include <atomic>

static std::atomic<int> g_stream_cb{0}, g_float_cb{0}, g_bg_cb{0};
static std::atomic<void*> g_last_ptr{nullptr};

Gpu2HostNotifier* Gpu2HostNotifier::gpu2hostNotifier = nullptr;
Gpu2HostNotifier::Gpu2HostNotifier(QObject* parent) : QObject(parent) {}
Gpu2HostNotifier::~Gpu2HostNotifier() {}
Gpu2HostNotifier* Gpu2HostNotifier::getInstance(QObject* parent) {
	if (!gpu2hostNotifier) gpu2hostNotifier = new Gpu2HostNotifier(parent);
	return gpu2hostNotifier;
}
void CUDART_CB Gpu2HostNotifier::dh2StreamingCallback(void* p) { g_last_ptr = p; ++g_stream_cb; }
void CUDART_CB Gpu2HostNotifier::dh2FloatStreamingCallback(void* p) { g_last_ptr = p; ++g_float_cb; }
void CUDART_CB Gpu2HostNotifier::backgroundSignalCallback(void* p) { g_last_ptr = p; ++g_bg_cb; }
void CUDART_CB Gpu2HostNotifier::bscanDisblayBufferReadySignalCallback(void*) {}
void CUDART_CB Gpu2HostNotifier::enfaceDisplayBufferReadySignalCallback(void*) {}
void CUDART_CB Gpu2HostNotifier::volumeDisblayBufferReadySignalCallback(void*) {}

extern "C" void refcuda_callback_counts(int* streaming, int* floatStreaming, int* background) {
	if (streaming) *streaming = g_stream_cb.load();
	if (floatStreaming) *floatStreaming = g_float_cb.load();
	if (background) *background = g_bg_cb.load();
}
