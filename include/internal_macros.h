/* Error-check macros of the plugin library; same names and behaviour as the reference's
 * stereoDNN/lib/internal_macros.h:14-37 (log through ILogger at kERROR with file:line:func, then assert). */
#ifndef REDTAIL_INTERNAL_MACROS_H
#define REDTAIL_INTERNAL_MACROS_H

#undef CHECKL
#define CHECKL(status, log)                                                                  \
    do {                                                                                     \
        auto res_ = (status);                                                                \
        if ((int)res_ != 0) redtail::tensorrt::reportError(res_, __FILE__, __LINE__, __FUNCTION__, log); \
    } while (false)

#undef CHECK
#define CHECK(status) CHECKL(status, log_)

#undef UNUSED
#define UNUSED(x) ((void)(x))

#undef UNUSEDR
#ifdef NDEBUG
#define UNUSEDR(x) ((void)(x))
#else
#define UNUSEDR(x)
#endif

#endif
