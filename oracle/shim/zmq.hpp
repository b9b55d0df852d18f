// STUB of cppzmq (zmq.hpp, ZeroMQ 4.3.x C++ binding) for compiling the reference's translation units in
// place: types and methods only, no transport.  Every call aborts -- the oracle harness never reaches one.
#pragma once
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <string>
#define ZMQ_DEALER 5
#define ZMQ_ROUTER 6
#define ZMQ_REQ 3
#define ZMQ_REP 4
#define ZMQ_PUB 1
#define ZMQ_SUB 2
#define ZMQ_PUSH 8
#define ZMQ_PULL 7
#define ZMQ_SNDMORE 2
#define ZMQ_DONTWAIT 1
#define ZMQ_NOBLOCK 1
#define ZMQ_IDENTITY 5
#define ZMQ_SUBSCRIBE 6
#define ZMQ_BACKLOG 19
#define ZMQ_SNDHWM 23
#define ZMQ_RCVHWM 24
#define ZMQ_LINGER 17
#define ZMQ_RCVTIMEO 27
#define ZMQ_SNDTIMEO 28
#define ZMQ_ROUTER_MANDATORY 33
#define ZMQ_POLLIN 1
#define ZMQ_IO_THREADS 1
typedef struct { void *socket; int fd; short events; short revents; } zmq_pollitem_t;
namespace zmq {
[[noreturn]] inline void stub_called() { std::abort(); }
struct error_t : std::exception { int num() const { return 0; } const char *what() const noexcept override { return "zmq stub"; } };
class message_t {
public:
    message_t() {}
    explicit message_t(size_t n) : n_(n), p_(n ? std::malloc(n) : nullptr) {}
    message_t(const void *src, size_t n) : n_(n), p_(n ? std::malloc(n) : nullptr) { if (n) std::memcpy(p_, src, n); }
    message_t(message_t &&o) noexcept : n_(o.n_), p_(o.p_) { o.p_ = nullptr; o.n_ = 0; }
    message_t &operator=(message_t &&o) noexcept { std::free(p_); n_ = o.n_; p_ = o.p_; o.p_ = nullptr; o.n_ = 0; return *this; }
    message_t(const message_t &) = delete;
    ~message_t() { std::free(p_); }
    void *data() { return p_; }
    const void *data() const { return p_; }
    size_t size() const { return n_; }
    void rebuild(size_t n) { std::free(p_); n_ = n; p_ = n ? std::malloc(n) : nullptr; }
    void rebuild() { rebuild(0); }
    void copy(const message_t *o) { rebuild(o->n_); if (n_) std::memcpy(p_, o->p_, n_); }
    void move(message_t *o) { *this = std::move(*o); }
    bool more() const { return false; }
private:
    size_t n_ = 0;
    void *p_ = nullptr;
};
class context_t {
public:
    context_t() {}
    explicit context_t(int) {}
    context_t(int, int) {}
    void close() {}
};
class socket_t {
public:
    socket_t(context_t &, int) {}
    socket_t(socket_t &&) {}
    void bind(const char *) { stub_called(); }
    void bind(const std::string &) { stub_called(); }
    void connect(const char *) { stub_called(); }
    void connect(const std::string &) { stub_called(); }
    void setsockopt(int, const void *, size_t) {}
    template <class T> void setsockopt(int, const T &) {}
    bool send(message_t &, int = 0) { stub_called(); }
    size_t send(const void *, size_t, int = 0) { stub_called(); }
    bool recv(message_t *, int = 0) { stub_called(); }
    size_t recv(void *, size_t, int = 0) { stub_called(); }
    void close() {}
    operator void *() { return nullptr; }
};
inline int poll(zmq_pollitem_t *, size_t, long = -1) { stub_called(); }
inline int poll(zmq_pollitem_t const *, size_t, long = -1) { stub_called(); }
}  // namespace zmq
