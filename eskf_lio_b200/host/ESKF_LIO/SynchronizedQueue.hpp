// SynchronizedQueue.hpp — the mutex queue the sensor callbacks fill
// (include/ESKF_LIO/SynchronizedQueue.hpp of the reference: push / pop /
// popAll).  Threading glue, kept only so Odometry has the reference's
// constructor; a deque under a lock_guard.
#ifndef ESKF_LIO_B200_SYNCHRONIZED_QUEUE_HPP_
#define ESKF_LIO_B200_SYNCHRONIZED_QUEUE_HPP_

#include <deque>
#include <mutex>
#include <optional>
#include <utility>

namespace ESKF_LIO
{
template<typename T>
class SynchronizedQueue
{
public:
  void push(T data)
  {
    std::lock_guard<std::mutex> lock(mutex_);
    items_.push_back(std::move(data));
  }

  std::optional<T> pop()
  {
    std::lock_guard<std::mutex> lock(mutex_);
    if (items_.empty()) {return std::nullopt;}
    std::optional<T> out(std::move(items_.front()));
    items_.pop_front();
    return out;
  }

  std::deque<T> popAll()
  {
    std::lock_guard<std::mutex> lock(mutex_);
    std::deque<T> out;
    out.swap(items_);
    return out;
  }

private:
  std::deque<T> items_;
  std::mutex mutex_;
};
}  // namespace ESKF_LIO

#endif  // ESKF_LIO_B200_SYNCHRONIZED_QUEUE_HPP_
