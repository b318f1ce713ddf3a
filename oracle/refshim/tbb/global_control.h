// Stand-in for <tbb/global_control.h> (see README.md): the thread cap is accepted and ignored (everything is serial).
#pragma once
#include <cstddef>
namespace tbb {
class global_control {
public:
    enum parameter { max_allowed_parallelism, thread_stack_size };
    global_control(parameter, std::size_t) {}
};
}  // namespace tbb
