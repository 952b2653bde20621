extern "C" {
size_t tsasr_joint_bwd_workspace_bytes(int, int, int, int, int, long long) { return 0; }
int tsasr_joint_bwd(const void*, const void*, const void*, const float*, const int32_t*, const int32_t*, const int32_t*,
                    int, int, int, int, int, int, int, float, const float*, const float*, const float*, const float*,
                    const float*, const float*, void*, size_t, long long, float*, float*, float*, float*, tsasr_stream_t) {
    return fail(TSASR_E_UNSUPPORTED, "not built yet");
}
}
