#include <stdio.h>
#include <fcntl.h>
#include <errno.h>
#include <string.h>
#include <unistd.h>
#include <sys/mount.h>
int main(){ int fd=open("/dev/fuse",O_RDWR); if(fd<0){printf("open /dev/fuse: %s\n",strerror(errno));return 1;}
 char opts[128]; snprintf(opts,sizeof opts,"fd=%d,rootmode=40000,user_id=0,group_id=0,allow_other",fd);
 int r=mount("fzfs","/tmp/fzmnt","fuse",MS_NOSUID|MS_NODEV,opts); printf("mount: %d %s\n",r,r?strerror(errno):"ok");
 if(!r){ umount2("/tmp/fzmnt",MNT_DETACH); } return 0; }
