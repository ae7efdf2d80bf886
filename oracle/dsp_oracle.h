/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  See dsp_oracle_impl.h. */
#ifndef DSP_ORACLE_H
#define DSP_ORACLE_H
#include <stdint.h>

/* DSPFatal conditions of the reference, in the order they are cited in
 * dsp_oracle_impl.h.  The numeric values are shared with the product's
 * row_status codes (include/dspeed_b200.h) so tests can compare them. */
enum {
  ORC_OK = 0,
  ORC_FATAL_PZ_NAN = 1,            /* pole_zero.py:76-77 */
  ORC_FATAL_DPZ_SHORT = 2,         /* pole_zero.py:145-148 */
  ORC_FATAL_RISE_NEG = 3,          /* trap_filters.py:53-54 */
  ORC_FATAL_FLAT_NEG = 4,          /* trap_filters.py:56-57 */
  ORC_FATAL_FALL_NEG = 5,          /* trap_filters.py:205-206 */
  ORC_FATAL_TRAP_WIDE = 6,         /* trap_filters.py:59-60 */
  ORC_FATAL_PICKOFF_NONINT = 7,    /* trap_filters.py:278-279 */
  ORC_FATAL_MW_RANGE = 8,          /* moving_windows.py:52-55 */
  ORC_FATAL_MWM_LEN_NONINT = 9,    /* moving_windows.py:167-168 */
  ORC_FATAL_MWM_NUM_NONINT = 10,   /* moving_windows.py:170-171 */
  ORC_FATAL_MWM_RANGE = 11,        /* moving_windows.py:173-174 */
  ORC_FATAL_MWM_NUM_NEG = 12,      /* moving_windows.py:176-177 */
  ORC_FATAL_TSTART_NONINT = 13,    /* time_point_thresh.py:67-68 */
  ORC_FATAL_WALK_NONINT = 14,      /* time_point_thresh.py:70-71 */
  ORC_FATAL_TSTART_RANGE = 15,     /* time_point_thresh.py:73-74 */
  ORC_FATAL_INTERP_MODE = 16,      /* time_point_thresh.py:222, fixed_time_pickoff.py:125 */
  ORC_FATAL_POLARITY_ZERO = 17,    /* time_point_thresh.py:314 */
  ORC_FATAL_FTP_INT = 18,          /* fixed_time_pickoff.py:85 */
  ORC_FATAL_WINDOWER_LEN = 19,     /* windower.py:42-43 */
  ORC_FATAL_UPSAMPLE = 20,         /* upsampler.py:41-42 */
  ORC_FATAL_CONV_KERNEL_LONG = 21, /* convolutions.py:48-49 */
  ORC_FATAL_CONV_MODE = 22,        /* convolutions.py:70 */
  ORC_FATAL_CONV_OUTLEN = 23,      /* convolutions.py:53-68 */
  ORC_FATAL_GMLE_LEN = 24,         /* get_multi_local_extrema.py:126-129 */
  ORC_FATAL_GMLE_DELTA = 25,       /* get_multi_local_extrema.py:130-131 */
  ORC_FATAL_GMLE_DIR = 26,         /* get_multi_local_extrema.py:305-306 */
  ORC_FATAL_RF_B_SCALAR = 27,      /* recursive_filter.py:66-67 */
  ORC_FATAL_RF_SHORT = 28,         /* recursive_filter.py:68-71 */
  ORC_FATAL_SHAPE = 29             /* gufunc shape mismatch (numpy raises) */
};
#endif
