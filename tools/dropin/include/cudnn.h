// sample_app/main.cpp:12 includes <cudnn.h> without using anything from it; the product does not depend on cuDNN, so the
// drop-in build resolves the include to this empty header instead of the system one.
#pragma once
