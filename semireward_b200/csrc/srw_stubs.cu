// srw_stubs — intentionally empty: every symbol declared in include/srw.h has a kernel behind it.
