void dav1d_log(void *c, const char *fmt, ...) { (void)c; (void)fmt; }
