/* Interface declarations for the subset of libuv 1.48 that bd-iaas-us/infiniStore uses.
 *
 * The image has no libuv development package, but the real library is present: uvloop 0.22.1
 * ships libuv 1.48.0 inside uvloop/loop.cpython-312-x86_64-linux-gnu.so and exports every
 * uv_* symbol.  The reference is linked against that shared object, so the event loop it
 * runs on IS libuv - this header only declares the functions and gives the handle / request
 * types their public leading fields and their exact sizes (queried from the library itself
 * with uv_handle_size() / uv_req_size(), checked again at start-up by baseline/refshim_check).
 */
#ifndef REFSHIM_UV_H
#define REFSHIM_UV_H

#include <netinet/in.h>
#include <semaphore.h>
#include <stddef.h>
#include <stdint.h>
#include <sys/socket.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UV_EOF (-4095)

typedef struct uv_loop_s uv_loop_t;
typedef struct uv_handle_s uv_handle_t;
typedef struct uv_stream_s uv_stream_t;
typedef struct uv_tcp_s uv_tcp_t;
typedef struct uv_poll_s uv_poll_t;
typedef struct uv_req_s uv_req_t;
typedef struct uv_write_s uv_write_t;
typedef struct uv_work_s uv_work_t;
typedef sem_t uv_sem_t;

typedef struct uv_buf_t {
    char* base;
    size_t len;
} uv_buf_t;

typedef void (*uv_alloc_cb)(uv_handle_t* handle, size_t suggested_size, uv_buf_t* buf);
typedef void (*uv_read_cb)(uv_stream_t* stream, ssize_t nread, const uv_buf_t* buf);
typedef void (*uv_write_cb)(uv_write_t* req, int status);
typedef void (*uv_connection_cb)(uv_stream_t* server, int status);
typedef void (*uv_close_cb)(uv_handle_t* handle);
typedef void (*uv_poll_cb)(uv_poll_t* handle, int status, int events);
typedef void (*uv_work_cb)(uv_work_t* req);
typedef void (*uv_after_work_cb)(uv_work_t* req, int status);

enum uv_poll_event { UV_READABLE = 1, UV_WRITABLE = 2, UV_DISCONNECT = 4, UV_PRIORITIZED = 8 };

/* Handles: `void* data; uv_loop_t* loop; uv_handle_type type;` lead every handle (public
 * fields of UV_HANDLE_FIELDS); the rest is private to the library. */
#define REFSHIM_UV_HANDLE(NAME, SIZE)                 \
    struct NAME {                                     \
        void* data;                                   \
        uv_loop_t* loop;                              \
        int type;                                     \
        char refshim_private_[(SIZE) - 2 * sizeof(void*) - sizeof(int)]; \
    }
REFSHIM_UV_HANDLE(uv_handle_s, 96);
REFSHIM_UV_HANDLE(uv_stream_s, 248);
REFSHIM_UV_HANDLE(uv_tcp_s, 248);
REFSHIM_UV_HANDLE(uv_poll_s, 160);

/* Requests: UV_REQ_FIELDS = `void* data; uv_req_type type; void* reserved[6];` (64 bytes). */
struct uv_req_s {
    void* data;
    int type;
    void* reserved[6];
};
struct uv_write_s {
    void* data;
    int type;
    void* reserved[6];
    uv_write_cb cb;
    uv_stream_t* send_handle;
    uv_stream_t* handle;
    char refshim_private_[192 - 64 - 3 * sizeof(void*)];
};
struct uv_work_s {
    void* data;
    int type;
    void* reserved[6];
    uv_loop_t* loop;
    uv_work_cb work_cb;
    uv_after_work_cb after_work_cb;
    char refshim_private_[128 - 64 - 3 * sizeof(void*)];
};

uv_loop_t* uv_default_loop(void);
const char* uv_strerror(int err);
const char* uv_err_name(int err);
uv_buf_t uv_buf_init(char* base, unsigned int len);
int uv_ip4_addr(const char* ip, int port, struct sockaddr_in* addr);

int uv_tcp_init(uv_loop_t*, uv_tcp_t* handle);
int uv_tcp_bind(uv_tcp_t* handle, const struct sockaddr* addr, unsigned int flags);
int uv_listen(uv_stream_t* stream, int backlog, uv_connection_cb cb);
int uv_accept(uv_stream_t* server, uv_stream_t* client);
int uv_read_start(uv_stream_t*, uv_alloc_cb alloc_cb, uv_read_cb read_cb);
int uv_write(uv_write_t* req, uv_stream_t* handle, const uv_buf_t bufs[], unsigned int nbufs,
             uv_write_cb cb);
void uv_close(uv_handle_t* handle, uv_close_cb close_cb);

int uv_poll_init(uv_loop_t* loop, uv_poll_t* handle, int fd);
int uv_poll_start(uv_poll_t* handle, int events, uv_poll_cb cb);
int uv_poll_stop(uv_poll_t* handle);

int uv_queue_work(uv_loop_t* loop, uv_work_t* req, uv_work_cb work_cb,
                  uv_after_work_cb after_work_cb);

int uv_sem_init(uv_sem_t* sem, unsigned int value);
void uv_sem_destroy(uv_sem_t* sem);
void uv_sem_post(uv_sem_t* sem);
void uv_sem_wait(uv_sem_t* sem);

size_t uv_handle_size(int type);
size_t uv_req_size(int type);
const char* uv_version_string(void);

#ifdef __cplusplus
}
#endif
#endif
