#ifndef CVSHIM_MF_TIMESYNC_H
#define CVSHIM_MF_TIMESYNC_H
#include "subscriber.h"
namespace message_filters {
template <class A, class B, class C>
class TimeSynchronizer {
  public:
    TimeSynchronizer(Subscriber<A>&, Subscriber<B>&, Subscriber<C>&, int) {}
    template <class F>
    void registerCallback(const F&)
    {
    }
};
} // namespace message_filters
#endif
