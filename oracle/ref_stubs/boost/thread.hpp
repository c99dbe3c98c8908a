// ORACLE — TEST INFRASTRUCTURE ONLY: the reference's solve polls boost's interruption point; no-op here.
#pragma once
namespace boost { namespace this_thread { inline void interruption_point() {} } struct thread_interrupted {}; }
