/*
 * dspeed_b200 -- entry points of a GENERATED chain kernel library (dspeed_b200/_chains/chain_<hash>.so).
 *
 * The chain compiler (dspeed_b200/codegen.py: specialised CTA-per-waveform kernel; dspeed_b200/warpchain.py:
 * warp-per-waveform kernel for short waveforms) turns one compiled ProcessingChain -- the processor list that the
 * reference executes block by block in ProcessingChain._execute_procs (src/dspeed/processing_chain.py:1144-1163) --
 * into ONE sm_100a kernel, compiles it with nvcc into its own shared library and loads it with ctypes.  Every such
 * library exports the same three C symbols; this header is the contract a host binding relies on.
 */
#ifndef DSPEED_B200_CHAIN_H
#define DSPEED_B200_CHAIN_H

#ifdef __cplusplus
extern "C" {
#endif

/* dynamic shared memory of the kernel in bytes (<= 227 KB) */
int chain_smem_bytes(void);

/* number of program nodes + 1 (size of the per-node cycle counters of the tracing build) */
int chain_n_nodes(void);

/*
 * Launch the chain over `n_rows` waveforms, asynchronously on `stream` (a cudaStream_t).
 *
 *   ptrs    host array of 2 * n_ptrs + 1 machine words:
 *             ptrs[0 .. n_ptrs)            DEVICE pointers of the input / output columns in the order the chain
 *                                          compiler assigned them (SpecChain.ptrs; inputs: waveform values, baseline,
 *                                          t0, dt ...; outputs: one column per requested scalar),
 *             ptrs[n_ptrs .. 2 n_ptrs)     row strides of those columns in ELEMENTS (int64),
 *             ptrs[2 n_ptrs]               index of the first row inside the caller's table (int64; reported with a
 *                                          data-dependent DSPFatal like processing_chain.py:1156-1159)
 *   fatal   device int32[4 * n_processors] fatal records (see dspeed_b200.h), may be NULL
 *   prof    device int64 counters of the tracing build, NULL in production
 *   num_sms number of persistent CTAs (one per SM)
 *
 * Returns 0, DSPB_ERR_UNSUPPORTED (pointer count / alignment), or -cudaError_t.  The library never allocates, frees
 * or retains the buffers; re-entrant across streams and devices.
 */
int chain_launch(const void* const* ptrs, long long n_ptrs, long long n_rows, int* fatal, long long* prof, int num_sms,
                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DSPEED_B200_CHAIN_H */
