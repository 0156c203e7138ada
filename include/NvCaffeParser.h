// The reference's sample_app/main.cpp:5 includes this header but uses nothing from it.
